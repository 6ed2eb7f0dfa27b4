python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r02_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo bench c2 rc=$?
python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo bench c3 rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo bench ref rc=$?
TAG=c2 ARGS="3600 16 44100" bash scripts/profile_r02.sh
TAG=c3_1h ARGS="3600 24 96000" bash scripts/profile_r02.sh
TAG=c3 LIGHT=1 ARGS="36000 24 96000" bash scripts/profile_r02.sh
