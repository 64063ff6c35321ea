cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -k "pipelined or slices" 2>&1 | tail -5
for s in 2 4; do T2D_FWD_SLICES=$s timeout 300 python tools/e2e_breakdown.py 65536 f32 2>&1 | grep "E=" | sed 's/.*total/total/'; done
