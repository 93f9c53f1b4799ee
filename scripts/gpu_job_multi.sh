#!/bin/bash
# multi-GPU bench exactly as the driver launches it: one rank per GPU over NCCL
N=${1:-2}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/m_build.log 2>&1
nvidia-smi -L > gpurun_out/m_smi.txt
for n in ${NS:-1 $N}; do
  if [ $n -eq 1 ]; then
    timeout -k 10 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err
  else
    timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/m_bench_n$n.json 2> gpurun_out/m_bench_n$n.err
    timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/m_bench_ref_n$n.json 2>> gpurun_out/m_bench_n$n.err
  fi
  echo "== n=$n rc=$?"; cat gpurun_out/m_bench_n$n.json; tail -3 gpurun_out/m_bench_n$n.err
done
