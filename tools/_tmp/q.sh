#!/bin/bash
pts=$1; shift
env "$@" EDGEFEM_B200_CLUSTER_PROF=1 python tools/cluster_probe.py --points $pts --reps 2 2>/tmp/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$pts', '$*', [('%.2f'%r['ms_kernel']) for r in d['runs']], d['runs'][-1]['shape'], d['runs'][-1]['rhs_iterations'])"
grep "cluster prof" /tmp/err.txt | tail -22
