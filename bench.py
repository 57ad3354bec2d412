#!/usr/bin/env python
"""bench.py — SCoNe train trajectories/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg5|cfg4|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one optimizer step of the 3-layer SCoNe (hidden 32) over one batch of synthetic trajectories on the
named complex: sparse flows -> fused layer forwards -> readout/NLL -> fused layer backwards -> (all-reduce of
the flat [grads | nll | count] buffer when N > 1) -> Adam.  Every trajectory in the batch contributes to the
loss (mask all ones).  Weak scaling: the per-GPU batch is fixed and the complex is replicated.

  value  whole-job trajectories/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e    same metric through the public host API (SconeModel.loss_grad + adam_step on pinned HOST buffers,
         H2D of the batch and D2H of the loss inside the timed region)
  roofline   the dominant kernel family, timed live with CUDA events on the launching stream; bytes follow the
             device-counted rows the flagged unit kernels produce (support of the trajectories)
  other_mode the same step with the dense zero-fill switched the other way (dense-stream mode: zero_fill_kernel at
             ~0.9 of the measured HBM peak is then the dominant kernel)
  cpu_baseline  the oracle's sparse CPU port (oracle/scone_oracle.py) on a bounded sample, rank 0, N = 1 only
--impl reference times that same CPU port as the whole arm (the reference's dense E x E formulation cannot be
instantiated at E = 1M: 4 TB per operator; jax itself is not installable offline — see DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n_nodes for the generator, per-GPU batch, hidden, micro-batch)
    'cfg5': dict(n_nodes=370000, batch=4096, hidden=32, micro_batch=4096,
                 desc='1M-edge synthetic holed Delaunay complex, 4096 trajectories per GPU (32768 at 8 GPUs), 3-layer SCoNe hidden 32'),
    'cfg4': dict(n_nodes=110000, batch=4096, hidden=32, micro_batch=4096,
                 desc='~300k-edge synthetic complex, batch 4096, 3-layer SCoNe hidden 32'),
    'cfg1': dict(n_nodes=400, batch=1000, hidden=16, micro_batch=1000,
                 desc='default synthetic complex (400 nodes), 1000 trajectories, 3-layer SCoNe hidden 16'),
}
KIND_NAMES = ['layer_fwd', 'layer_bwd', 'layer0_fwd', 'layer0_bwd', 'readout', 'flows_to_dense', 'zero_fill', 'cone']


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  The timed region of
    this bench is tens of milliseconds, shorter than one `nvidia-smi` invocation, so the sampler polls NVML in-process
    (nvidia_ml_py, ~1 kHz, own thread); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.rows, self.proc, self.index, self.nv, self.h = [], None, index, None, None
        self.running, self.th, self.source = False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(',')[self.index])
                except Exception:
                    idx = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.running, self.source = True, 'nvml'
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = 'nvidia-smi'
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = [(getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8), 'hw_slowdown'),
                (getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40), 'hw_thermal_slowdown'),
                (getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), 'sw_thermal_slowdown'),
                (getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4), 'sw_power_cap')]
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(self.h))
                self.rows.append([sm, self.mx, 0.0] + ['Active' if r & bit else 'Not Active' for bit, _ in bits])
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.nv is not None:
            self.running = False
            self.th.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        else:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no NVML, no nvidia-smi'], 'samples': 0}
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(self.NAMES, r[3:7]):
                    if str(v).lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        load = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {'sm_mhz': (load[len(load) // 2] if load else None), 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm), 'source': self.source}


def ncu_traffic(kernel, E, mb, C, zero_fill, pipeline=3):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set full`
    captures -- only reported when this run has the same launch shape as the capture (cfg5).
      zero_fill  profiles/prof_fill_r1m_raw.csv       (dense-stream mode, b = 64)
      layer_fwd / layer_bwd  profiles/prof_cone_r1z9_raw.csv   (cone pipeline, b = 4096: the two layer_fwd_rows_kernel /
                                                       rows_bwd_kernel launches of one step, averaged)"""
    src, pat = None, None
    if kernel == 'zero_fill' and (E, mb, C) == (999308, 64, 32):
        src = 'prof_fill_r1m_raw.csv'
    if kernel in ('layer_fwd', 'layer_bwd') and not zero_fill and pipeline == 3 and (E, mb, C) == (999308, 4096, 32):
        src, pat = 'prof_cone_r1z9_raw.csv', {'layer_fwd': 'layer_fwd_rows_kernel', 'layer_bwd': 'rows_bwd_kernel'}[kernel]
    if src is None:
        return None
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, 'profiles', src))))
        hdr, unit = rows[0], rows[1]
        kn = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tot, n = 0.0, 0
        for val in rows[2:]:
            if len(val) < len(hdr) or (pat and (kn is None or pat not in val[kn])):
                continue
            for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                i = hdr.index(name)
                tot += float(val[i]) * scale[unit[i]]
            n += 1
        return tot / n if n else None
    except Exception:
        return None


def make_dataset(cfg, rank):
    from scone_gcn_b200 import synthetic_data_gen as sdg
    if cfg['n_nodes'] <= 1000:
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        from golden_util import Dataset
        ds = Dataset('dataset_default.npz')
        sp = sdg.SparseDataset.from_dense(ds.flows, ds.B1, ds.B2, ds.targets, ds.train_mask, ds.test_mask, ds.last_nodes,
                                          ds.target_nodes)
        return sp
    return sdg.generate_sparse_dataset(cfg['n_nodes'], cfg['batch'], seed=1030 + rank, n_waypoints=24)


def tri_lists(sp):
    from scone_gcn_b200.complex import incidence_lists_from_simplices
    return incidence_lists_from_simplices(sp.edges, sp.faces)


def cpu_port_run(sp, hidden, n_sample, repeats, seed=0):
    """Oracle port (CPU, torch sparse, all host threads): fwd + bwd over `n_sample` trajectories; trajectories/s."""
    import numpy as np
    import torch
    from oracle import scone_oracle as so
    en, es, te, ts = tri_lists(sp)
    orc = so.SparseOracle('scone', sp.edges, te, ts, int(sp.n_nodes))
    rs = np.random.RandomState(1030)
    shapes = [(1, hidden)] * 3 + [(hidden, hidden)] * 6 + [(hidden, 1)]
    W = [0.01 * rs.randn(*s) for s in shapes]
    E = len(sp.edges)
    X = np.zeros((E, n_sample), np.float32)
    for t in range(n_sample):
        sl = slice(sp.traj_ptr[t], sp.traj_ptr[t + 1])
        X[sp.flow_edge[sl], t] = sp.flow_val[sl]
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.loss_and_grads(W, X, sp.last_nodes[:n_sample], sp.target_idx[:n_sample], np.ones(n_sample, np.float32))
        times.append(time.perf_counter() - t0)
    return n_sample / min(times), torch.get_num_threads(), times


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    import torch
    sp = make_dataset(cfg, 0)
    n_sample = 2 if cfg['n_nodes'] > 200000 else (8 if cfg['n_nodes'] > 1000 else 64)
    t0 = time.perf_counter()
    tps, threads, times = cpu_port_run(sp, cfg['hidden'], n_sample, args.warmup + args.steps)
    times = times[args.warmup:] or times
    tps = n_sample / (sum(times) / len(times))
    sample = '%d trajectories fwd+bwd per step on the same complex (E=%d), sparse CSR CPU port of the reference maths' % (
        n_sample, len(sp.edges))
    out = {'impl': 'reference', 'metric': 'SCoNe train trajectories/sec', 'value': tps, 'unit': 'trajectories/s',
           'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(times) / len(times),
           'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': args.config + ': ' + cfg['desc'], 'E': int(len(sp.edges)), 'N': int(sp.n_nodes),
                      'F': int(len(sp.faces)), 'sample_trajectories_per_step': n_sample},
           'cpu_baseline': {'value': tps, 'unit': 'trajectories/s', 'cores': threads, 'kind': 'port', 'sample': sample},
           'e2e': {'value': tps, 'unit': 'trajectories/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0,
           'note': 'jax is not installable offline and the dense E x E reference formulation cannot be instantiated at this '
                   'size; this is the oracle port (oracle/scone_oracle.py SparseOracle) on torch CPU sparse kernels'}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg5', choices=sorted(CONFIGS))
    ap.add_argument('--batch', type=int, default=0, help='override the per-GPU batch')
    ap.add_argument('--micro-batch', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=20)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', dest='extras', action='store_false', help='skip the sparse-mode and dense-kernel extras')
    ap.add_argument('--zero-fill', type=int, default=0, help='timed region with dense zero-fill of the outputs on (1) or off (0)')
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg['batch'] = args.batch
    if args.micro_batch:
        cfg['micro_batch'] = args.micro_batch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, cfg, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import scone_gcn_b200 as sg
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU path)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ['NCCL_DEBUG_FILE'] = '/dev/stderr'       # keep stdout to the one JSON line (NCCL prints its version banner)
        dist.init_process_group('nccl', device_id=dev)
    L = sg.lib()

    t_setup = time.time()
    sp = make_dataset(cfg, rank)
    B = sp.n_traj if args.config == 'cfg1' else cfg['batch']
    cx = sg.SimplicialComplex.from_simplices(int(sp.n_nodes), sp.edges, sp.faces, 'scone')
    C, mb = cfg['hidden'], min(cfg['micro_batch'], B)
    net = sg.SconeModel(cx, [C, C, C], micro_batch=mb)
    rs = np.random.RandomState(1030)                       # same init on every rank
    net.set_weights([0.01 * rs.randn(*s) for s in net.shapes])
    E, N, F, D = cx.E, cx.N, cx.F, cx.D
    nnz = int(sp.traj_ptr[B])
    # host batch in pinned memory (e2e) and its device copy (value)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    h = dict(ptr=pinned(sp.traj_ptr[:B + 1].astype(np.int32)), edge=pinned(sp.flow_edge[:nnz].astype(np.int32)),
             val=pinned(sp.flow_val[:nnz].astype(np.float32)), last=pinned(sp.last_nodes[:B].astype(np.int32)),
             tgt=pinned(sp.target_idx[:B].astype(np.int32)), mask=pinned(np.ones(B, np.float32)))
    d = {k: v.to(dev) for k, v in h.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in h.values())
    stream = torch.cuda.current_stream().cuda_stream
    gbuf = net.grads_tensor()
    from scone_gcn_b200 import _lib
    lr, wd = 1e-3, 5e-5
    step_no = [0]

    def step_dev():
        _lib.check(L.scone_model_loss_grad_dev(net.handle, B, _lib.dptr(d['ptr']), _lib.dptr(d['edge']), _lib.dptr(d['val']),
                                               _lib.dptr(d['last']), _lib.dptr(d['tgt']), _lib.dptr(d['mask']), 1, stream))
        if world > 1:
            dist.all_reduce(gbuf)
        net.adam_step(step_no[0], lr, wd, stream)
        step_no[0] += 1

    hp = {k: v.numpy() for k, v in h.items()}
    loss_host = np.zeros(net.n_params + 2, np.float32)

    def step_e2e():
        net.loss_grad(hp['ptr'], hp['edge'], hp['val'], hp['last'], hp['tgt'], hp['mask'], zero_first=True, stream=stream, read=False)
        if world > 1:
            dist.all_reduce(gbuf)
        net.adam_step(step_no[0], lr, wd, stream)
        step_no[0] += 1
        buf = net.read_grads(stream)                       # D2H of [grads | nll_sum | count]: the step's loss
        return float(buf[-2] / buf[-1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    net.set_zero_fill(args.zero_fill)
    for _ in range(args.warmup):
        step_dev()
    barrier()
    setup_s = time.time() - t_setup

    L.scone_profile_reset()
    L.scone_profile_enable(1)
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = L.scone_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_dev()
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = L.scone_launch_count() - launches0
    L.scone_profile_enable(0)
    clk = clocks.stop()
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # per-kernel-family device time inside the timed region
    import ctypes

    def read_families():
        fam_ = {}
        for k, nme in enumerate(KIND_NAMES):
            n_l, t_ms = ctypes.c_int64(), ctypes.c_double()
            L.scone_profile_read(k, n_l, t_ms)
            fam_[nme] = (n_l.value, t_ms.value)
        rows_ = {}
        for k, nme in ((0, 'layer_fwd'), (1, 'layer_bwd')):
            r = ctypes.c_int64()
            L.scone_profile_read_rows(k, r)
            rows_[nme] = r.value
        return fam_, rows_
    fam, rows = read_families()
    # ALGORITHMIC bytes.  The flagged unit kernels produce only the (edge, trajectory) rows inside the
    # trajectories' support (counted on the device): per produced row a fused layer forward reads one input row and
    # writes one output row (4*(Cin+Cout) bytes), a backward reads G and Hin rows and writes Gprev (4*(2*Cin+Cout));
    # neighbour re-reads are cache-served, exactly as in the dense formula of SURVEY 8(d).  zero_fill (dense-stream mode
    # only) writes 4*E*b*C bytes per launch.
    peak, peak_src = load_peaks()

    def kernel_table(fam_, rows_):
        alg_ = {'zero_fill': 4.0 * E * mb * C}
        if fam_['layer_fwd'][0]:
            alg_['layer_fwd'] = rows_['layer_fwd'] * 4.0 * 2 * C / fam_['layer_fwd'][0]
        if fam_['layer_bwd'][0]:
            alg_['layer_bwd'] = rows_['layer_bwd'] * 4.0 * 3 * C / fam_['layer_bwd'][0]
        tot_ms_ = sum(t for _, t in fam_.values()) or 1.0
        ks = {}
        for nme, (n_l, t_ms) in fam_.items():
            if n_l:
                avg = t_ms / n_l
                ks[nme] = {'launches': n_l, 'avg_ms': avg, 'share_of_kernel_time': t_ms / tot_ms_}
                if nme in alg_:
                    ks[nme].update({'algorithmic_bytes_per_launch': alg_[nme], 'achieved_gbs': alg_[nme] / avg / 1e6,
                                    'frac': alg_[nme] / avg / 1e6 / peak})
        return ks
    kernels = kernel_table(fam, rows)
    # the roofline is quoted for the tensor-moving family with the largest share (the families with a byte model); `cone` is the
    # bit-level set-up of the row lists (integer work, latency-bound: cone / live-row marking, compaction, clearing)
    largest = max(kernels, key=lambda k_: kernels[k_]['share_of_kernel_time'])
    with_bytes = [k_ for k_ in kernels if 'achieved_gbs' in kernels[k_]]
    dom = max(with_bytes or kernels, key=lambda k_: kernels[k_]['share_of_kernel_time'])
    pipeline_id = L.scone_model_get_pipeline(net.handle)
    dense_bytes_per_traj = 4.0 * E * (15 * C + 2)
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kernels[dom].get('achieved_gbs'), 'peak': peak, 'unit': 'GB/s',
                'frac': kernels[dom].get('frac'), 'traffic': ncu_traffic(dom, E, mb, C, args.zero_fill, pipeline_id), 'peak_source': peak_src,
                'largest_family': largest,
                'algorithmic_bytes_per_launch': kernels[dom].get('algorithmic_bytes_per_launch'),
                'rows_per_step': {k_: v_ / args.steps for k_, v_ in rows.items()},
                'dense_rows_per_step': float(E) * B * 2,
                'kernels': kernels,
                'dense_equivalent': {'bytes_per_trajectory': dense_bytes_per_traj,
                                     'effective_gbs': dense_bytes_per_traj * value / world / 1e9,
                                     'note': 'SURVEY 8(d) dense formula 4*E*(15C+2) bytes per trajectory times the measured per-GPU '
                                             'trajectories/s: what a dense-streaming implementation would have to move to match'},
                'bytes_model': 'layer_fwd / layer_bwd: rows produced (device-counted) x 4*(Cin+Cout) / 4*(2*Cin+Cout) bytes per launch; the backward '
                               'family time includes the weight-gradient GEMM and its reduction; cone = bit-level set-up of the row lists '
                               '(receptive cone, live rows, compaction, clearing). The row kernels are latency / issue bound (16 warps per SM, '
                               'dependent entry -> bitmap -> row loads; profiles/prof_cone_r1z9_*), not HBM bound; zero_fill (dense-stream '
                               'mode): 4*E*b*C bytes per launch',
                'pipeline': {3: 'row lists over the readout cone, compact tensors', 2: 'row lists, compact tensors', 1: 'row lists, dense tensors', 0: 'unit kernels, byte flags'}[
                    pipeline_id]}

    # end to end through the host API
    barrier()
    e2e_steps = max(1, args.e2e_steps)
    step_e2e()
    barrier()
    ev0.record()
    loss = None
    for _ in range(e2e_steps):
        loss = step_e2e()
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e = {'value': world * B * e2e_steps / (e2e_ms / 1e3), 'unit': 'trajectories/s', 'h2d_bytes_per_step': h2d_bytes,
           'd2h_bytes_per_step': int(4 * (net.n_params + 2)), 'steps': e2e_steps, 'last_loss': loss}

    # extra 1: the same step in the OTHER mode (results bit-identical, tests/): dense-stream mode bulk-zeroes every dense
    # [E][b][C] tensor once per micro-batch (zero_fill_kernel on a side stream is then the HBM-bound dominant kernel)
    other_mode = None
    if args.extras:
        mb2 = min(64, B)                                    # dense tensors: 6 x E x mb2 x C x 4 bytes must fit
        net2 = sg.SconeModel(cx, [C, C, C], micro_batch=mb2, zero_fill=not args.zero_fill)
        net2.set_weights(net.get_weights())

        def step_other():
            _lib.check(L.scone_model_loss_grad_dev(net2.handle, B, _lib.dptr(d['ptr']), _lib.dptr(d['edge']), _lib.dptr(d['val']),
                                                   _lib.dptr(d['last']), _lib.dptr(d['tgt']), _lib.dptr(d['mask']), 1, stream))
            net2.adam_step(step_no[0], lr, wd, stream)
        step_other()
        barrier()
        L.scone_profile_reset()
        L.scone_profile_enable(1)
        other_steps = min(args.steps, 5)
        ev0.record()
        for _ in range(other_steps):
            step_other()
        ev1.record()
        barrier()
        L.scone_profile_enable(0)
        sm_ms = max_over_ranks(ev0.elapsed_time(ev1))
        fam2, rows2 = read_families()
        mb_main, mb = mb, mb2
        k2 = kernel_table(fam2, rows2)
        mb = mb_main
        other_mode = {'zero_fill': 1 - args.zero_fill, 'micro_batch': mb2, 'value': world * B * other_steps / (sm_ms / 1e3),
                      'unit': 'trajectories/s', 'ms_per_step': sm_ms / other_steps, 'steps': other_steps, 'zero_fill_kernel': k2.get('zero_fill'),
                      'note': 'zero_fill=1 (dense-stream mode, unit kernels): every activation / gradient tensor is a complete dense array '
                              '(each byte written once per micro-batch by zero_fill_kernel, timed on its side stream); zero_fill=0: rows '
                              'outside the support are never written'}
        del net2

    # extra 2: the contracted DENSE-tile measurement (north-star / SURVEY 8d): one fused 32->32 layer on dense random
    # features, no occupancy information, algorithmic bytes 4*E*b*(Cin+Cout) fwd and 4*E*b*(2*Cout+Cin) bwd
    roofline_dense = None
    if args.extras and rank == 0:
        bd = min(mb, 64)
        Hd = torch.randn(E, bd, C, device=dev)
        Od = torch.empty_like(Hd)
        Wd = [torch.randn(C, C, device=dev) * 0.2 for _ in range(3)]
        wsd = torch.empty(L.scone_layer_backward_workspace_bytes(C, C) // 4 + 16, device=dev)
        dWd = torch.zeros(3, C, C, device=dev)

        def t_of(fn, it=3):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(it):
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                z.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(z))
            return best
        f_ms = t_of(lambda: _lib.check(L.scone_layer_forward(cx.handle, 0, bd, C, C, _lib.dptr(Hd), _lib.dptr(Wd[0]), _lib.dptr(Wd[1]),
                                                             _lib.dptr(Wd[2]), _lib.dptr(Od), None, None, None, stream)))
        b_ms = t_of(lambda: _lib.check(L.scone_layer_backward(cx.handle, 0, bd, C, C, _lib.dptr(Hd), _lib.dptr(Od), _lib.dptr(Wd[0]),
                                                              _lib.dptr(Wd[1]), _lib.dptr(Wd[2]), _lib.dptr(Od), _lib.dptr(dWd), 0,
                                                              _lib.dptr(wsd), None, None, None, None, stream)))
        fa, ba = 4.0 * E * bd * 2 * C / f_ms / 1e6, 4.0 * E * bd * 3 * C / b_ms / 1e6
        roofline_dense = {'bound': 'hbm', 'unit': 'GB/s', 'peak': peak, 'b': bd,
                          'layer_fwd': {'ms': f_ms, 'achieved': fa, 'frac': fa / peak},
                          'layer_bwd': {'ms': b_ms, 'achieved': ba, 'frac': ba / peak},
                          'note': 'one fused 32->32 layer on dense random features, no flags (every row computed): forward = slab kernel '
                                  '(merged-row gather into mma.sync fragments, 3xTF32 product), backward = fp32 SIMT tile kernel; '
                                  'algorithmic bytes 4*E*b*(Cin+Cout) / 4*E*b*(2*Cout+Cin); not in the timed region'}
        del Hd, Od

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sample = 2 if E > 500000 else (8 if E > 2000 else 64)
        tps, threads, times = cpu_port_run(sp, C, n_sample, 2)
        cpu_baseline = {'value': tps, 'unit': 'trajectories/s', 'cores': threads, 'kind': 'port',
                        'sample': '%d trajectories fwd+bwd on the same complex (E=%d), best of 2; sparse CSR CPU port '
                                  '(oracle/scone_oracle.py), torch CPU sparse kernels' % (n_sample, E)}
    if rank == 0:
        # what one step touches (bytes): with the compact pipelines the tensors follow the rows actually produced
        n_mb = (B + mb - 1) // mb
        if pipeline_id >= 2 and not args.zero_fill:
            fr, br = rows['layer_fwd'] / args.steps, rows['layer_bwd'] / args.steps
            touched = (2 * 3 * 4.0 * E * mb / 1024 * n_mb          # summary words of the 2L two-level bitmaps, scanned
                       + fr * 4.0 * C * 2 + br * 4.0 * C * 5      # produced rows: H and G rows, A_k rows of the weight-gradient GEMM
                       + 8.0 * nnz)                                # flows
            l2_policy = ('no explicit flush; one step touches ~%.0f MB of bitmaps summaries, compact tensor rows and flows (plus the operator '
                         'rows it walks), %s the 126 MB L2; the dense address space of the same tensors is %.0f GB'
                         % (touched / 1e6, 'more than' if touched > 126e6 else 'LESS than (this workload is L2-resident by size)',
                            4.0 * E * B * C * 6 / 1e9))
        else:
            l2_policy = ('inputs larger than L2: the activation / gradient tensors of one micro-batch span %.2f GB and every step walks %d '
                         'micro-batches' % (4.0 * E * mb * C * 6 / 1e9, n_mb))
        out = {'metric': 'SCoNe train trajectories/sec', 'value': value, 'unit': 'trajectories/s', 'n_gpus': world,
               'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
               'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
               'config': {'workload': args.config + ': ' + cfg['desc'], 'N': N, 'E': E, 'F': F, 'D': D,
                          'per_gpu_batch': B, 'global_batch': B * world, 'hidden': C, 'layers': 3, 'micro_batch': mb,
                          'zero_fill': args.zero_fill,
                          'parallelism': 'dp%d (trajectory shards, complex replicated, one all-reduce of %d floats per step)'
                                         % (world, net.n_params + 2),
                          'l2_policy': l2_policy,
                          'generator_seed': 1030, 'mean_flow_nnz': nnz / B},
               'roofline': roofline, 'roofline_dense': roofline_dense, 'other_mode': other_mode, 'e2e': e2e,
               'cpu_baseline': cpu_baseline, 'gpu_launches': int(launches),
               'clocks': clk, 'setup_s': setup_s}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
