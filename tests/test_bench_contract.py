"""CPU checks of the measurement contract: the committed bench line of the final build (profiles/r02zz_bench.json, written by
`python bench.py` on a B200) carries every key the driver reads, and the reference arm / CLI surface of bench.py is intact."""
import json
import os
import subprocess
import sys

import parity

PROFILES = os.path.join(parity.ROOT, 'profiles')


def _line(name):
    with open(os.path.join(PROFILES, name)) as fh:
        return json.load(fh)


def test_committed_bench_line_has_the_contract_keys():
    for name, n_gpu in (('r02zz_bench.json', 1), ('r02zz_bench_n8.json', 8)):
        d = _line(name)
        for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                  'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
            assert k in d, (name, k)
        assert d['n_gpus'] == n_gpu and d['higher_is_better'] is True and d['vs_baseline'] is None and d['data'] == 'synthetic'
        assert 'workload' in d['config'] and 'model' not in d['config']
        assert d['gpu_launches'] > 0
        for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
            assert k in d['e2e'], k
        assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
        assert 0 < d['e2e']['value'] < d['value']                      # end to end includes the copies
        for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
            assert k in d['roofline'], k
        assert abs(d['roofline']['frac'] - d['roofline']['achieved'] / d['roofline']['peak']) < 1e-6
        for k in ('value', 'unit', 'cores', 'kind', 'sample'):
            assert k in d['cpu_baseline'], k
        assert d['cpu_baseline']['kind'] == 'reference'
        for k in ('sm_mhz', 'sm_max_mhz', 'reasons'):
            assert k in d['clocks'], k
        assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    one, eight = _line('r02zz_bench.json'), _line('r02zz_bench_n8.json')
    assert one['scaling'] == 'strong' and eight['scaling'] == 'strong'     # config 3 = 4096 replicas in total
    assert eight['weak_scaling']['scaling'] == 'weak'
    for blk in ('config1', 'config4', 'config5', 'single_replica_latency'):
        assert blk in one and one[blk].get('cpu_baseline'), blk


def test_traffic_json_is_stamped():
    t = _line('traffic.json')
    assert 'source_hash' in t and len(t['source_hash']) == 16
    sys.path.insert(0, parity.ROOT)
    import bench
    # bench.py quotes ncu numbers only for the build they were captured from
    assert isinstance(bench.kernel_source_hash(), str) and len(bench.kernel_source_hash()) == 16


def test_bench_cli_surface():
    r = subprocess.run([sys.executable, os.path.join(parity.ROOT, 'bench.py'), '--help'], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    for flag in ('--gpus', '--steps', '--warmup', '--impl'):
        assert flag in r.stdout + r.stderr      # (bench.py keeps stdout for its one JSON line)


def test_replica_sharding_covers_every_replica_once():
    """config 3 is 4096 replicas IN TOTAL: rank r of N gets a contiguous block, blocks tile the range"""
    sys.path.insert(0, parity.ROOT)
    import bench
    for total, world in ((4096, 1), (4096, 2), (4096, 8), (256, 8), (48, 8), (10, 4)):
        seen = []
        for rank in range(world):
            first, n = bench._shard(total, world, rank)
            seen.extend(range(first, first + n))
        assert seen == list(range(total)), (total, world)
