#!/bin/bash
pts=$1; shift
env "$@" EDGEFEM_B200_SMALL_PROF=1 python tools/cluster_probe.py --points $pts --reps 4 2>/tmp/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['runs'][-1]; print('$pts', '$*', [('%.2f'%r['ms_kernel']) for r in d['runs']], r['shape'], r['converged'], r['rhs_iterations'])"
grep "small prof" /tmp/err.txt | tail -1 | sed 's/.*cycles: //' | cut -c1-200
