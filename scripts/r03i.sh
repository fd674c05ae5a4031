set -x
mkdir -p gpurun_out/r03i
N=${NGPU:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 400 $T bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r03i/bench$N.json 2> gpurun_out/r03i/bench$N.err; grep "^{" gpurun_out/r03i/bench$N.json | cut -c1-400; tail -5 gpurun_out/r03i/bench$N.err | cut -c1-300
