for kb in 1024 256 4096 65536; do
GCNN_PACKED_PIECE_KB=$kb python scripts/ab_host_paths.py > gpurun_out/r2zd_ab_host_$kb.json 2> gpurun_out/r2zd_ab_host.err; tail -3 gpurun_out/r2zd_ab_host.err; echo piece $kb; cat gpurun_out/r2zd_ab_host_$kb.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:v for k,v in d.items() if k.startswith((\"e2e\",\"config1\"))})"
done
