out=gpurun_out/r4c; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x -k "meyer or golden or ragged or c1 or config or efficient or binding" > $out/pytest.log 2>&1; tail -4 $out/pytest.log
timeout 300 python tools/bench_configs.py c1 c2 c5:28 > $out/configs.jsonl 2> $out/configs.err; cat $out/configs.jsonl | cut -c1-700; tail -3 $out/configs.err
PQC_MW=generic timeout 300 python tools/bench_configs.py c1 c2 > $out/configs_generic_mw.jsonl 2>> $out/configs.err; cut -c1-500 $out/configs_generic_mw.jsonl
