cd $GRAFT_REPO_ROOT
for sl in 2 8,12,16 10,16 8,13,16 9,14,16; do echo "slices $sl"; T2D_CHUNKS=16 T2D_FWD_SLICES=$sl timeout 300 python tools/e2e_breakdown.py 65536 f32 2>&1 | grep "E=" | sed 's/.*total/total/'; done
for sl in 2 8,12,16 ; do echo "u8 slices $sl"; T2D_CHUNKS=16 T2D_FWD_SLICES=$sl timeout 300 python tools/e2e_breakdown.py 65536 u8 2>&1 | grep "E=" | sed 's/.*total/total/'; done
