#!/bin/bash
pts=$1; shift
env "$@" python tools/cluster_probe.py --points $pts --reps 4 2>/tmp/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$pts', '$*', [('%.2f'%r['ms_kernel']) for r in d['runs']], d['runs'][-1]['shape'], d['runs'][-1]['converged'])"
