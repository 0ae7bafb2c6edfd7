#!/bin/bash
# Programmatic dependent launch: between the persistent per-layer kernels (MPDB_PDL_LAYERS, default 1) and between the cluster-kernel
# forwards and single-wave guide launches of the small-batch loop (MPDB_PDL_LOOP, default 1) vs plain graph edges (0).
for pdl in 0 1; do for wl in cfg4 cfg2 cfg4_ddim; do MPDB_PDL_LOOP=$pdl python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl_loop=$pdl $wl', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['forward_us_by_precision'])"; done; done
for pdl in 0 1; do for wl in cfg5 cfg3; do MPDB_PDL_LAYERS=$pdl python bench.py --workload $wl --steps 9 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl_layers=$pdl $wl', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['forward_us_by_precision'])"; done; done
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -4
