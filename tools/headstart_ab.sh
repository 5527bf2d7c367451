for i in 1 2; do
for v in "" 1; do echo "NO_HEADSTART=$v: $(EVAC_BENCH_NO_HEADSTART=$v python bench.py --steps 2000 --warmup 100 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step']*1e3, d['l2_resident']['ms_per_step']*1e3, d['rollout']['ms_per_step']*1e3, d['clocks'])")"; done; done
