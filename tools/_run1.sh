timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_all4.log 2>&1; tail -3 gpurun_out/t_all4.log
b() { timeout 300 python bench.py --no-cpu-baseline 2>> gpurun_out/bench_t.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'])"; }
b "streams c10"
S2VT_WAVE_CHUNK=6 b "streams c6"
S2VT_BWD_SERVER_CTAS=16 S2VT_FWD_SERVER_CTAS=20 b "streams c10 ctas 20/16"
timeout 300 python tools/timeline_step.py > gpurun_out/timeline_streams.txt 2>&1
