"""bench.py — headline benchmark of the B200 hot paths (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

metric  : sampled edges/s of pyg_lib.sampler.neighbor_sample on BASELINE.json configs[1]
          (ogbn-products-shaped CSR, 2,449,029 nodes / 123,718,280 edges, int64, fan-out [15,10],
          1024 seeds per step, no replacement, return_edge_id).  One step = one call on one batch.
value   : whole-job edges/s with the graph AND the seeds resident in HBM (each call still ends with the
          API's own host sync, because the op returns Python lists).
e2e     : same metric through the public API with HOST buffers: pinned seeds -> device every step and
          the sampled (row, col, node_id, edge_id) copied back to pinned host memory every step.
The same JSON line carries `roofline` (dominant kernel, device-timed), `cpu_baseline` (the reference's
CPU implementation timed on this box's host cores, bounded sample) and `segment_matmul` (BASELINE.json
configs[2]: 64 relations, N=2^20 ragged rows, 128->128 bf16, TFLOP/s + roofline fractions).
N>1: every rank holds a replica of the CSR and samples its own disjoint seed batches (weak scaling;
no data-path collective) — the frontier-sharded all-gather variant is pyg_lib_b200.sampler.dist.
"""
import argparse
import ctypes as C
import json
import os
import os.path as osp
import statistics
import subprocess
import sys
import threading
import time

ROOT = osp.dirname(osp.abspath(__file__))
for p in (ROOT, osp.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

N_NODES, N_EDGES = 2_449_029, 123_718_280
BATCH, FANOUT = 1024, [15, 10]
BYTES_PER_EDGE = 41.0  # algorithmic bytes per sampled edge at fan-out [15,10], int64


class ClockMonitor:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.lo, self.hi = index, None, [], 0, None

    def start(self):
        """Launch nvidia-smi and wait for its first sample: NVML start-up takes ~100 ms and holds driver locks,
        so it must not overlap the timed region."""
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append(ln)
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:
                time.sleep(0.01)
        except Exception:  # noqa
            self.proc = None
        return self

    def __enter__(self):   # marks the start of the timed region
        self.lo = len(self.lines)
        return self

    def __exit__(self, *a):  # marks its end
        time.sleep(0.06)       # one more sample so that short regions are covered
        self.hi = len(self.lines)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines[max(self.lo - 1, 0):self.hi]:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def run_ref_bench(args):
    out = subprocess.run([sys.executable, osp.join(ROOT, 'oracle', 'ref_bench.py')] + args, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=3000)
    for ln in out.stdout.splitlines():
        if ln.startswith('REFBENCH '):
            return json.loads(ln[len('REFBENCH '):])
    raise RuntimeError('ref_bench failed: ' + out.stderr[-2000:])


def reference_arm(a):
    """--impl reference: the reference's CPU implementation on the host cores, same metric/config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, cores)  # every host thread: one single-threaded worker process per core (how PyG deploys it)
    per_step_calls = 2  # each "step" = a bounded sample: `workers` processes x 2 calls of 1024 seeds
    total_calls = per_step_calls * (a.steps + a.warmup)
    total_calls = max(2, min(total_calls, 40))  # keep the whole run within a few minutes
    r = run_ref_bench(['sampler', '--workers', str(workers), '--calls', str(total_calls)])
    line = {'impl': 'reference', 'metric': 'sampled_edges_per_s', 'value': r['edges_per_s'], 'unit': 'edges/s',
            'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': 1e3 * r['seconds'] / max(r['calls'] / workers, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64', 'data': 'synthetic', 'config': config_dict(a.gpus),
            'cpu_baseline': {'value': r['edges_per_s'], 'unit': 'edges/s', 'cores': r['cores'], 'kind': r['kind'],
                             'sample': r['sample']},
            'e2e': {'value': r['edges_per_s'], 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def config_dict(n_gpus):
    return {'workload': 'neighbor_sample ogbn-products-shaped CSR (2,449,029 nodes / 123,718,280 edges, int64, '
                        'log-normal degrees), fanout [15,10], 1024 seeds/step, replace=False, return_edge_id=True',
            'batch_seeds': BATCH, 'fanout': FANOUT, 'index_dtype': 'int64',
            'parallelism': 'replicas x%d (disjoint seed batches per rank, CSR replicated)' % n_gpus,
            'l2': 'inputs_larger_than_L2 (col = 990 MB, new random seeds every step)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-matmul', action='store_true')
    a = ap.parse_args()
    if a.impl == 'reference':
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    from graphs import lognormal_csr, ragged_ptr
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    import pyg_lib_b200 as P
    abi = C.CDLL(osp.join(osp.dirname(P.__file__), 'libpyg_b200.so'))

    # ---- synthetic inputs (deterministic; CPU-generated so the CPU baseline sees the same graph)
    rowptr_c, col_c = lognormal_csr(N_NODES, N_EDGES, seed=1)
    rowptr, col = rowptr_c.to(dev), col_c.to(dev)
    del col_c
    perm = torch.randperm(N_NODES, generator=torch.Generator().manual_seed(2))
    n_batches = N_NODES // BATCH
    my_batches = [(rank + i * world) % n_batches for i in range(a.steps + a.warmup)]
    seeds_host = [perm[b * BATCH:(b + 1) * BATCH].clone().pin_memory() for b in my_batches]
    seeds_dev = [s.to(dev) for s in seeds_host]
    torch.manual_seed(12345 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mon = ClockMonitor(local).start()

    def timed(step_fn, finish=None):
        for i in range(a.warmup):
            step_fn(i)
        if finish is not None:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = P.kernel_launches()
        edges = 0
        with mon:
            e0.record()
            for i in range(a.warmup, a.warmup + a.steps):
                edges += step_fn(i)
            if finish is not None:
                finish()   # e.g. make the timed stream wait for outstanding result copies
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = P.kernel_launches() - l0
        t = torch.tensor([ms, float(edges), float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, edges, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
        barrier()
        return ms, edges, int(launches), mon.summary()

    # ---- value: inputs resident in HBM
    def step_dev(i):
        out = P.sampler.neighbor_sample(rowptr, col, seeds_dev[i], FANOUT)
        return out[0].numel()
    ms, edges, launches, clocks = timed(step_dev)
    value = edges / (ms * 1e-3)

    # ---- e2e: host seeds in (pinned -> device) and host results out (device -> pinned) EVERY step, through the
    # public API.  Like a double-buffered loader, the result copy of step i runs on a copy stream while step
    # i+1 samples (the calls themselves cannot overlap: each consumes the CPU generator where the previous one
    # left it); the timed region ends only after the last copy has landed.
    cap = BATCH * (FANOUT[0] + FANOUT[0] * FANOUT[1]) + BATCH
    host_out = [[torch.empty(cap, dtype=torch.int64).pin_memory() for _ in range(4)] for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    d2h = [0]

    def step_e2e(i):
        s = seeds_host[i].to(dev, non_blocking=True)
        outs = P.sampler.neighbor_sample(rowptr, col, s, FANOUT)[:4]
        ready = torch.cuda.Event()
        ready.record()
        slot = i & 1
        copied[slot].synchronize()          # host buffers of step i-2 are free again
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            for h, t in zip(host_out[slot], outs):
                t.record_stream(copy_stream)
                h[:t.numel()].copy_(t, non_blocking=True)
            copied[slot].record()
        d2h[0] += 8 * sum(t.numel() for t in outs)
        return outs[0].numel()

    def finish_e2e():
        torch.cuda.current_stream().wait_stream(copy_stream)
    ms_e, edges_e, _, _ = timed(step_e2e, finish_e2e)
    # (d2h counter also ran during warm-up; per-step figure from the timed steps only)
    e2e = {'value': edges_e / (ms_e * 1e-3), 'unit': 'edges/s', 'h2d_bytes_per_step': BATCH * 8,
           'd2h_bytes_per_step': int(d2h[0] / max(a.steps + a.warmup, 1))}

    line = {'metric': 'sampled_edges_per_s', 'value': value, 'unit': 'edges/s', 'n_gpus': world, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'int64', 'data': 'synthetic', 'config': config_dict(world), 'clocks': clocks,
            'e2e': e2e, 'gpu_launches': launches, 'edges_per_step': edges / a.steps / world}

    if rank == 0:
        # ---- roofline of the dominant sampler kernel: CUDA events on the launching stream around every launch
        peaks = {}
        try:
            peaks = json.load(open(osp.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:  # noqa
            pass
        # DRAM traffic per launch of the dominant kernels, from the committed `ncu --set full` capture
        # (profiles/ncu_summary_r1.json: dram__bytes_read.sum + dram__bytes_write.sum, MB)
        traffic = {}
        try:
            summ = json.load(open(osp.join(ROOT, 'profiles', 'ncu_summary_r1.json')))
            for grp in summ.values():
                for d in grp:
                    nm = 'k_sample' if 'k_sample' in d['Kernel Name'] else ('k_segment_matmul_tc' if 'k_segment_matmul_tc' in d['Kernel Name'] else None)
                    if nm:
                        traffic.setdefault(nm, []).append(1e6 * (float(d['dram__bytes_read.sum']) + float(d['dram__bytes_write.sum'])))
            traffic = {k: sum(v) / len(v) for k, v in traffic.items()}
        except Exception:  # noqa
            traffic = {}
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s'
        abi.pygb200_profile_enable(1)
        n_prof = min(50, a.steps)
        for i in range(a.warmup, a.warmup + n_prof):
            step_dev(i)
        torch.cuda.synchronize()
        abi.pygb200_profile_enable(0)
        prof = {}
        for name in ('count', 'sample', 'mark', 'assign', 'lookup'):
            msv, ln, wk = C.c_double(), C.c_int64(), C.c_int64()
            abi.pygb200_profile_read(name.encode(), C.byref(msv), C.byref(ln), C.byref(wk))
            prof[name] = (msv.value, ln.value, wk.value)
        dom = max(prof, key=lambda k: prof[k][0])
        d_ms, d_launches, d_work = prof[dom]
        # the second-hop launch carries ~90% of the edges; report per-launch averages over both hops
        bytes_per_launch = BYTES_PER_EDGE * d_work / max(d_launches, 1)
        achieved = bytes_per_launch / (d_ms / max(d_launches, 1) * 1e-3) / 1e9 if d_ms > 0 else 0.0
        line['roofline'] = {'bound': 'hbm', 'kernel': 'k_' + dom + '_s' if dom != 'lookup' else 'k_lookup', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                            'frac': achieved / hbm_peak, 'traffic': traffic.get('k_' + dom), 'peak_source': peak_src,
                            'avg_launch_us': 1e3 * d_ms / max(d_launches, 1),
                            'bytes_per_launch': bytes_per_launch,
                            'kernel_ms_share': {k: v[0] for k, v in prof.items()},
                            'note': 'C2 is latency-bound (5 MB/call): the fraction is reported, the binding limit is '
                                    'the per-call critical path (launches + one D2H)'}

        # ---- segment_matmul (BASELINE configs[2]); single GPU
        if not a.no_matmul:
            Nn, K, M, B = 1 << 20, 128, 128, 64
            g = torch.Generator().manual_seed(0)
            x = torch.randn(Nn, K, generator=g).to(torch.bfloat16).to(dev)
            w = (torch.randn(B, K, M, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
            ptr = ragged_ptr(Nn, B, 100).to(dev)
            for _ in range(5):
                y = P.ops.segment_matmul(x, ptr, w)
            torch.cuda.synchronize()
            iters = 50
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                y = P.ops.segment_matmul(x, ptr, w)
            e1.record(); torch.cuda.synchronize()
            mm_ms = e0.elapsed_time(e1) / iters
            flops, byts = 2.0 * Nn * K * M, Nn * K * 2 + Nn * M * 2 + B * K * M * 2 + (B + 1) * 8
            tf_peak = float(peaks.get('bf16_tflops', 1590.0))
            line['segment_matmul'] = {
                'config': '64 relations, N=2^20 ragged rows (log-normal lengths, one empty), 128->128 bf16, ptr on device',
                'ms': mm_ms, 'tflops': flops / (mm_ms * 1e-3) / 1e12, 'algorithmic_bytes': byts,
                'roofline': {'bound': 'hbm', 'achieved': byts / (mm_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': byts / (mm_ms * 1e-3) / 1e9 / hbm_peak, 'traffic': traffic.get('k_segment_matmul_tc'),
                             'tensor_frac_of_bf16_peak': flops / (mm_ms * 1e-3) / 1e12 / tf_peak},
                'note': 'x (256 MiB) + out (256 MiB) > L2; arithmetic intensity 63.75 FLOP/B => HBM-bound'}
            del x, w, y

        # ---- CPU baseline: the reference's own CPU path on this box's host cores (bounded sample)
        if world == 1 and not a.no_cpu_baseline:
            try:
                r = run_ref_bench(['sampler', '--workers', '1', '--calls', '200'])
                line['cpu_baseline'] = {'value': r['edges_per_s'], 'unit': 'edges/s', 'cores': r['cores'], 'kind': r['kind'],
                                        'sample': r['sample'], 'host_cores_available': os.cpu_count()}
                if not a.no_matmul:
                    rm = run_ref_bench(['matmul', '--calls', '3'])
                    line['segment_matmul']['cpu_baseline'] = {'value': rm['tflops'], 'unit': 'TFLOP/s', 'cores': rm['cores'],
                                                              'kind': rm['kind'], 'sample': rm['sample']}
            except Exception as ex:  # noqa
                line['cpu_baseline'] = {'value': None, 'unit': 'edges/s', 'cores': 0, 'kind': 'failed', 'sample': str(ex)[:300]}
        print(json.dumps(line), flush=True)
    mon.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
