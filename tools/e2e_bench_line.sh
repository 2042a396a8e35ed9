#!/bin/bash
# prints the e2e part of one bench.py run: tools/e2e_bench_line.sh [bench.py args]
python bench.py "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench e2e', round(d['e2e']['value']/1e9,2), 'G samples/s; plain D2H limit', round(d['e2e']['limit_gbs'],1), 'GB/s; PCIe MB per step', d['e2e']['d2h_bytes_per_step']>>20)"
