for W in 4; do echo W$W; RVLP_LIB=$PWD/build_variants/lib_gps_w$W.so python tools/gp_smem_lat.py 120 2>&1 | tail -7 | sed -n '1p;3p;7p'; done
