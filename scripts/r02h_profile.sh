set -x
mkdir -p gpurun_out/r02h
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:conv_tc3_kernel -s 27 -c 2 -o gpurun_out/r02h/conv_tc3_fold python scripts/prof_emb.py emb 64 > gpurun_out/r02h/t1.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc4_kernel -s 8 -c 2 -o gpurun_out/r02h/conv_tc4 python scripts/prof_emb.py emb 64 > gpurun_out/r02h/t2.log 2>&1
$NCU --set full --import-source on -k regex:conv_block32_kernel -s 3 -c 1 -o gpurun_out/r02h/conv_block32 python scripts/prof_emb.py emb 64 > gpurun_out/r02h/t3.log 2>&1
ls -la gpurun_out/r02h
