SH='32,256,72,3;32,72,144,3;32,144,216,3;64,128,40,3;64,40,72,3;64,72,112,3;64,120,64,3;16,224,72,3;16,144,216,3;32,224,128,1'
for kb in 96 64 48 32 24; do echo "== STAGE_KB=$kb"; SEMB_TC_STAGE_KB=$kb python scripts/bench_layers.py --only conv --shapes "$SH" 2>&1 | grep -v wgrad; done
for bs in 2 4; do echo "== BSPLIT=$bs"; SEMB_TMA_BSPLIT=$bs python scripts/bench_layers.py --only conv --shapes "$SH" 2>&1 | grep -v wgrad; done
