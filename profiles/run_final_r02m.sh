#!/bin/bash
# final validation of the round: GPU suite (default and with a 64-record stage = the direct/pending path under stress), smoke, bench
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ISX_STAGE=64 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02m_cfg3_n1.json 2> gpurun_out/r02m_cfg3_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02m_cfg3_n1.json').read())
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['parity']['bit_exact'], 'frac', round(d['roofline']['frac'],3), d['clocks'], 'cpu', d['cpu_baseline']['value'] if d.get('cpu_baseline') else None, 'launches', d['gpu_launches'])"
