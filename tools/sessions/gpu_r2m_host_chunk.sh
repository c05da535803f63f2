for HC in 4096 8192; do
  export RRTMG_TUNE="host_chunk=$HC,run_chunk=$HC"
  echo "host_chunk=$HC"
  bash tools/gpu_scale.sh 8
  cp gpurun_out/scale_T170L60_8gpu.json gpurun_out/scale_T170L60_8gpu_hc$HC.json
done
