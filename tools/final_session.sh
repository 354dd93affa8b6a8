#!/bin/bash
# One single-GPU measurement session for the records under profiles/: bench lines, warm pass times, the ncu launch list of the bench
# command and the ncu --set full captures.  usage (on the GPU box): bash tools/final_session.sh <tag>
cd "$(dirname "$0")/.."
T=${1:-r3}
mkdir -p gpurun_out
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -c 600 gpurun_out/${T}_bench_default.json; echo
python bench.py --workload seq1080p --no-cpu-baseline > gpurun_out/${T}_bench_seq1080p.json 2> gpurun_out/${T}_bench_seq1080p.err
python bench.py --workload views256 --no-cpu-baseline > gpurun_out/${T}_bench_views256.json 2> gpurun_out/${T}_bench_views256.err
python tools/pass_times.py > gpurun_out/${T}_pass_times_1080p.txt 2>&1; cat gpurun_out/${T}_pass_times_1080p.txt
python tools/pass_times.py --width 3840 --height 2160 > gpurun_out/${T}_pass_times_4k.txt 2>&1; cat gpurun_out/${T}_pass_times_4k.txt
{ python tools/frame_time.py --frames 256; python tools/frame_time.py --frames 256 --txaa; python tools/frame_time.py --frames 256 --txaa --no-godrays; python tools/frame_time.py --frames 64 --width 3840 --height 2160 --txaa; } > gpurun_out/${T}_frame_times.txt 2>&1; cat gpurun_out/${T}_frame_times.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cloud_raymarch -s 1 -c 1 -f -o gpurun_out/prof_cloud_${T} python tools/profile_frame.py --passes cloud > gpurun_out/ncu_${T}.log 2>&1
ncu --set full --clock-control none --import-source on -s 24 -c 8 -f -o gpurun_out/prof_passes_${T} python tools/profile_frame.py --width 1920 --height 1080 --passes frame --reps 6 > gpurun_out/ncu_passes_${T}.log 2>&1
ls -la gpurun_out | grep ${T}
