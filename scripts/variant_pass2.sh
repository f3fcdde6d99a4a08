# experiment helper: time the polygon workloads with each library variant in gpurun_variants/ (see build_variants.sh)
for v in default $(ls gpurun_variants 2>/dev/null | sed 's/\.so$//'); do
  if [ $v = default ]; then unset MR_B200_LIB; else export MR_B200_LIB=$PWD/gpurun_variants/$v.so; fi
  echo "== $v"
  REPS=3 python scripts/profile_config5.py 100000 1 | tail -1
  REPS=3 python scripts/profile_config5.py 100000 2 | tail -1
  if [ "${WITH_SMALL:-0}" = 1 ]; then
    REPS=5 python scripts/profile_config5.py 100000 1 8 64 0 | tail -1
    REPS=5 python scripts/profile_config5.py 100000 0 8 64 0 | tail -1
  fi
done
