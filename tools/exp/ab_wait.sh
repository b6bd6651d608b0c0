for lib in "" _w500 _w0 "" _w500 _w0; do
  if [ -n "$lib" ]; then export LMPCR_B200_LIB=3d_multiview_reg_b200/liblmpcr_b200$lib.so; else unset LMPCR_B200_LIB; fi
  python tools/filter_bench.py --pairs 296 --iters 3 | head -1 | sed "s/^/lib$lib: /"
done
