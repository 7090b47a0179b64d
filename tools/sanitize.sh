#!/bin/bash
# compute-sanitizer racecheck + memcheck over every FIR kernel (SURVEY §5 "race detection").
# usage (on a GPU box): tools/sanitize.sh [tag]   -> gpurun_out/<tag>/sanitizer_*.log
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, env...
  local name=$1 tool=$2; shift 2
  env "$@" timeout 400 $CS --tool $tool --print-limit 20 python tools/sanitize_workload.py > $OUT/sanitizer_${tool}_${name}.log 2>&1
  echo "$tool $name rc=$? :: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' $OUT/sanitizer_${tool}_${name}.log | tail -1)"
}
run default racecheck ADT_SANITIZE_FAMILIES=p32
run default memcheck ADT_SANITIZE_FAMILIES=p32
run cluster_c2 racecheck ADT_SANITIZE_FAMILIES=c2 ADT_SANITIZE_EPILOGUE=0
run tma racecheck ADT_SANITIZE_FAMILIES=p32 ADT_SANITIZE_SIZES=4096,8192 ADT_FIR_TMA=1 ADT_SANITIZE_EPILOGUE=0
run vt2 racecheck ADT_SANITIZE_FAMILIES=p32 ADT_SANITIZE_SIZES=4096,8192 ADT_FIR_VT=2 ADT_SANITIZE_EPILOGUE=0
run persist racecheck ADT_SANITIZE_FAMILIES=p32 ADT_FIR_PERSIST=2 ADT_SANITIZE_EPILOGUE=0
if [ -f pyaudiodsptools_b200/libadt_b200_ab.so ]; then
  run ab_p16 racecheck ADT_LIB_PATH=$PWD/pyaudiodsptools_b200/libadt_b200_ab.so ADT_SANITIZE_FAMILIES=p16 ADT_SANITIZE_EPILOGUE=0
  run ab_persist racecheck ADT_LIB_PATH=$PWD/pyaudiodsptools_b200/libadt_b200_ab.so ADT_SANITIZE_FAMILIES=p32 ADT_FIR_PERSIST=2 ADT_SANITIZE_EPILOGUE=0
fi
