cd $GRAFT_REPO_ROOT
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
run() { tag=$1; shift; env "$@" timeout 200 ncu --metrics $M --clock-control none -k regex:tc_gemm_fused -s 1 -c 1 --csv --log-file gpurun_out/r02_fz_$tag.csv python tools/glm_eval_bench.py --reps 1 --check 0 > /dev/null 2>&1; echo "$tag $@"; grep -E "dram__bytes|gpu__time|hit_rate|tensor" gpurun_out/r02_fz_$tag.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}'; }
run s2r2 B2M_TC_FUSE_SLAB=2 B2M_TC_FUSE_RING=2
run s2r3 B2M_TC_FUSE_SLAB=2 B2M_TC_FUSE_RING=3
run s1r3 B2M_TC_FUSE_SLAB=1 B2M_TC_FUSE_RING=3
run s4r2 B2M_TC_FUSE_SLAB=4 B2M_TC_FUSE_RING=2
run s4r3h B2M_TC_FUSE_SLAB=4 B2M_TC_FUSE_RING=3 B2M_TC_FUSE_HINTS6=4
run s4r3n B2M_TC_FUSE_SLAB=4 B2M_TC_FUSE_RING=3 B2M_TC_FUSE_HINTS5=0
