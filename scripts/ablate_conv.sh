#!/bin/bash
# Ablation of the conv pipeline stages on three HBM-bound layer shapes (profiling aid: SEMB_TC_DEBUG bits 1 no A loads,
# 2 no MMAs, 4 no stores, 8 no moments; SEMB_TC_PER_SM caps CTAs per SM; SEMB_TMA_STAGES ring depth)
SHAPES="256,8,8,3;256,32,16,3;256,32,32,1"
run() { echo "== $*"; env "$@" python scripts/bench_layers.py --only conv --shapes "$SHAPES" 2>&1 | grep -v wgrad; }
run SEMB_TC_DEBUG=0
run SEMB_TC_DEBUG=1
run SEMB_TC_DEBUG=2
run SEMB_TC_DEBUG=4
run SEMB_TC_DEBUG=7
run SEMB_TC_DEBUG=6
run SEMB_TMA_STAGES=2
run SEMB_TMA_STAGES=3
run SEMB_TC_PER_SM=1
run SEMB_TC_PER_SM=1 SEMB_TC_DEBUG=1
