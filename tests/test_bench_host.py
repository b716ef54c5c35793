# -*- coding: utf-8 -*-
"""
Host-side pieces of bench.py that need no GPU: the clock sampler's filtering of nvidia-smi samples by timestamp, the
choice of the CPU arm (the unmodified reference from baseline/_ref when present, else the oracle port), the
flop-equivalent counts of SURVEY.md 8(d) and the traffic figures read from profiles/.
"""
import datetime
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _line(t, sm, power, reasons=('Not Active',) * 4):
    stamp = datetime.datetime.fromtimestamp(t).strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]
    return f'{stamp}, 0, {sm}, 1965, {power}, 0x0000000000000000, ' + ', '.join(reasons)


class _Done:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0


def _sampler_with(lines, begin, end):
    s = bench.ClockSampler(0)
    s.file.write('\n'.join(lines) + '\n')
    s.proc = _Done()
    s.t_begin, s.t_end = begin, end
    return s


def test_clock_sampler_reports_the_samples_inside_the_timed_region():
    t0 = 1_800_000_000.0
    lines = [_line(t0 + 0.01 * i, 1200 if i < 10 else 1965, 150.0 if i < 10 else 320.0) for i in range(30)]
    lines[25] = _line(t0 + 0.25, 1500, 330.0, ('Not Active', 'Not Active', 'Not Active', 'Active'))   # after the region
    out = _sampler_with(lines, t0 + 0.12, t0 + 0.20).stop()
    assert out['sampled'] == 'inside the timed region'
    assert 8 <= out['samples'] <= 11 and out['samples_total'] == 30
    assert out['sm_mhz'] == 1965.0 and out['sm_max_mhz'] == 1965.0 and out['reasons'] == []
    # a throttle reason inside the region is reported
    lines[15] = _line(t0 + 0.15, 1700, 900.0, ('Not Active', 'Active', 'Not Active', 'Active'))
    out = _sampler_with(lines, t0 + 0.12, t0 + 0.20).stop()
    assert out['reasons'] == ['hw_thermal_slowdown', 'sw_power_cap']


def test_clock_sampler_falls_back_to_the_samples_under_load():
    t0 = 1_800_000_000.0
    lines = [_line(t0 + 0.01 * i, 1965, 300.0 if 5 <= i < 9 else 90.0) for i in range(12)]
    out = _sampler_with(lines, t0 + 5.0, t0 + 5.001).stop()           # a region no sample fell into
    assert out['samples'] == 4 and out['sampled'].startswith('under load next to')
    assert _sampler_with([], t0, t0 + 1).stop()['samples'] == 0
    assert _sampler_with(['garbage', '1, 2'], t0, t0 + 1).stop()['sm_mhz'] is None


def test_cpu_arm_prefers_the_reference_install(monkeypatch):
    have = os.path.isfile(os.path.join(bench.REF_DIR, 'xicsrt', '__init__.py'))
    monkeypatch.setattr(bench, '_REF', {})
    monkeypatch.setenv('XRT_BENCH_CPU_PORT', '1')
    assert bench.reference_module() is None and bench.cpu_kind() == 'port'
    monkeypatch.setattr(bench, '_REF', {})
    monkeypatch.delenv('XRT_BENCH_CPU_PORT')
    mod = bench.reference_module()
    if have:
        assert mod is not None and bench.cpu_kind() == 'reference'
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(bench.REF_DIR))
        assert not any(os.path.realpath(p) == os.path.realpath(bench.REF_DIR) for p in sys.path)
    else:
        assert mod is None and bench.cpu_kind() == 'port'


def test_flop_equivalents_follow_the_survey_rule():
    # F(cfg1) = 133 + 25 + 12 + 13 + f_b 78 + f_r 51 ~ 225 at the reference fractions
    assert abs(bench.flops_config2(0.523, 0.0128) - (183.0 + 0.523 * 78.0 + 0.0128 * 51.0)) < 1e-12
    assert 224.0 < bench.flops_config2(0.523, 0.0128) < 225.5
    f3, layers = bench.flops_config3(0.525, 0.1646)
    assert 1.0 < layers < 15.0 and f3 > bench.flops_config2(0.525, 0.1646)
    assert bench.flops_config5(0.5, 0.01) == bench.flops_config2(0.5, 0.01) + 60.0
    assert bench.flops_config4(0.51, 0.51) == 133.0 + 50.0 * 32 + 0.51 * (400.0 + 20.0 + 480.0 + 13.0) + 0.51 * 51.0


def test_traffic_file_matches_the_committed_digests():
    prof = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
    step = prof['step']['dram_bytes_per_launch']
    assert step == sum(prof[k][f] for k in ('k_cull32', 'k_trace') for f in ('dram_bytes_read', 'dram_bytes_write'))
    for name, c in prof['configs'].items():
        assert c['dram_bytes_per_launch'] == sum(k['dram_bytes_read'] + k['dram_bytes_write'] for k in c['kernels'].values()), name

    def digest(path):
        unit = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        vals = {}
        for line in open(os.path.join(ROOT, 'profiles', path)):
            f = line.strip().split(',')
            if len(f) == 3 and f[0].startswith('dram__bytes'):
                vals[f[0]] = float(f[2]) * unit[f[1]]
        return vals
    d = digest('r02_ncu_c2_cull_digest.csv')
    assert abs(d['dram__bytes_read.sum'] - prof['k_cull32']['dram_bytes_read']) < 1e3
    assert abs(d['dram__bytes_write.sum'] - prof['k_cull32']['dram_bytes_write']) < 1e3
    d = digest('r02_ncu_c5_cull_digest.csv')
    assert abs(d['dram__bytes_read.sum'] - prof['configs']['config5']['kernels']['k_cull32<bundles>']['dram_bytes_read']) < 1e3
