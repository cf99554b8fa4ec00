#!/usr/bin/env python
"""Benchmark of the radar GNN hot path (BASELINE.json metric: radar frames/s and edges/s, GNN forward and
forward+backward, % of roofline, CPU reference beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--points P] [--impl reference]

A step = one pass of the hot path over one batch of synthetic RadarScenes-shaped frames:
graph construction (kNN + radius degree + features) -> node/edge encoders -> 7 message-passing layers ->
4 heads.  Workload at N=1 is BASELINE.json configs[1]: 256 accumulated frames x 3000 points (reference-default
symmetrised-kNN graph, k=10, eps^2=25).  Under torchrun every rank processes its own 256 frames (weak scaling,
no data-path collective: frames are independent graphs).

`--impl reference` times the reference's CPU algorithm (oracle port of graph_features.py + gnn_detector.py in
plain PyTorch/NumPy, the per-frame loop of Model_Training.forward) on the host cores, on a bounded sample
of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from graph_neural_network_for_radar_perception_b200 import synth  # noqa: E402

KNN, EPS2 = 10, 25
GRID_MAX_R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
GRID_MAX_TH = np.pi * 0.5
CKPT = os.path.join(ROOT, 'tests', 'golden', 'graph_based_detector.pt')


def make_frames(n_frames, n_points, seed0=0, distinct=16):
    """`distinct` different synthetic frames, tiled to n_frames (generation is host-side NumPy and not timed)."""
    base = [synth.make_frame(seed0 + i, n_points, knn=KNN) for i in range(min(distinct, n_frames))]
    return [base[i % len(base)] for i in range(n_frames)]


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = [int(s[0]) for s in self.samples if len(s) >= 6 and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) >= 6 and s[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v == 'Active'})
        return {'sm_mhz': int(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference algorithm)
# -------------------------------------------------------------------------------------------------
def cpu_frame_pass(sd, data, lab_src, train=False, times=None):
    """One frame through the oracle port of the reference's per-frame path (graph_features.py:58-164 ->
    gnn_detector.py:141-162 [-> loss.py:37-76 + autograd when train]).  times (optional list of 3) accumulates the
    seconds of the graph / forward / loss+backward legs (SURVEY.md section 8d asks for them separately)."""
    from oracle import graph_np, model_torch
    t0 = time.perf_counter()
    adj = graph_np.adjacency_information(data, EPS2, KNN)
    nf = torch.from_numpy(graph_np.node_features(data, adj['degree'], True, 0, GRID_MAX_R, 0, GRID_MAX_TH).astype(np.float32))
    ef = torch.from_numpy(graph_np.edge_features(data, adj['adj_list']).astype(np.float32))
    ei = torch.from_numpy(adj['adj_list'])
    t1 = time.perf_counter()
    lab = synth.make_labels(data, lab_src, adj['adj_list'])
    cl = [torch.from_numpy(c) for c in lab['cluster_node_idx']]
    t2 = time.perf_counter()
    t3 = t2
    if train:
        sdg = {k: v.requires_grad_(True) for k, v in sd.items()}
        labels = {'cluster_node_idx': [cl], **{k: [torch.from_numpy(lab[k])] for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets')}}
        loss, _, _ = model_torch.training_forward(sdg, [nf], [ef], [ei], labels)
        t3 = time.perf_counter()
        sum(loss.values()).backward()
        for v in sdg.values():
            v.grad = None
    else:
        with torch.no_grad():
            model_torch.detector_forward(sd, nf, ef, ei, cl)
        t3 = time.perf_counter()
    t4 = time.perf_counter()
    if times is not None:
        times[0] += t1 - t0
        times[1] += t3 - t2
        times[2] += t4 - t3
    return ei.shape[1]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = torch.load(CKPT, map_location='cpu', weights_only=True)
    sample = max(1, min(args.ref_frames, args.frames))
    frames = make_frames(sample, args.points)
    for _ in range(args.warmup):
        cpu_frame_pass(sd, *frames[0])
    t0 = time.perf_counter()
    edges = 0
    for _ in range(args.steps):
        for f in frames:
            edges += cpu_frame_pass(sd, *f)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {'impl': 'reference', 'metric': 'radar_frames_per_s_gnn_fwd', 'value': fps, 'unit': 'frames/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args),
            'edges_per_s': edges / dt,
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{sample} frames x {args.points} points per step (graph build + forward, per-frame loop)'},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def workload_config(args):
    return {'workload': f'batched inference: {args.frames} synthetic accumulated frames x {args.points} points, '
                        f'symmetrised-kNN graph k={KNN} + radius degree eps2={EPS2} (BASELINE.json configs[1])',
            'frames_per_gpu': args.frames, 'points_per_frame': args.points, 'knn': KNN,
            'weights': 'reference checkpoint 1718175257362', 'cache': 'inputs larger than L2 (no flush needed)'}


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist
    from graph_neural_network_for_radar_perception_b200 import config, Model_Training, _cabi
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _cabi.lib()

    model = Model_Training(config(), dev)
    model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True))
    model = model.to(dev)
    det = model.pred.eval()

    frames = make_frames(args.frames, args.points, seed0=1000 * rank)
    # host buffers (pinned) of the raw accumulated points: what the reference's data loader hands over
    keys = ('meas_px', 'meas_py', 'meas_vx', 'meas_vy', 'meas_vr', 'meas_rcs', 'meas_timestamp')
    host = {}
    for k in keys:
        a = np.concatenate([f[0][k] for f in frames])
        host[k] = torch.from_numpy(a.astype(np.int64 if k == 'meas_timestamp' else np.float32)).pin_memory()
    fp = [0]
    for f in frames:
        fp.append(fp[-1] + f[0]['meas_px'].shape[0])
    n_nodes = fp[-1]
    # clusters: synthetic blobs + singletons (labels are inputs of the object head)
    pts_dev = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    bf0 = gf.build_graph_batch(pts_dev, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH)
    cl_lists = []
    for (d, src) in frames:
        blob = src['blob_of']
        cl = [np.nonzero(blob == b)[0] for b in np.unique(blob[blob >= 0])] + [np.array([i]) for i in np.nonzero(blob < 0)[0]]
        cl_lists.append([torch.from_numpy(c.astype(np.int64)) for c in cl])
    bf0.gb.set_clusters(cl_lists, fp[:-1], dev)
    cl_ptr, cl_members, n_clusters = bf0.gb.cl_ptr, bf0.gb.cl_members, bf0.gb.n_clusters
    n_edges, n_und = bf0.gb.n_edges, bf0.gb.n_und

    def device_step(pts):
        bf = gf.build_graph_batch(pts, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH)
        bf.gb.cl_ptr, bf.gb.cl_members, bf.gb.n_clusters = cl_ptr, cl_members, n_clusters
        with torch.no_grad():
            return det.forward_batch(bf.gb, bf.node_features, bf.edge_features, training=False)

    # End-to-end step: H2D of this step's inputs (pinned), graph build + forward, D2H of this step's results (pinned).
    # The read-back runs on a side stream behind an event, into one of two host buffer sets, so that it overlaps the next
    # step's compute (a serving loop hands results over one step late); every copy of every step lies inside the timed
    # region, which ends with a device-wide synchronize.
    out_host = [None, None]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_count = 0

    def e2e_step():
        nonlocal e2e_count
        slot = e2e_count & 1
        e2e_count += 1
        pts = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        outs = device_step(pts)
        if out_host[slot] is None:
            out_host[slot] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        else:
            copied[slot].synchronize()          # the host has consumed / may overwrite this buffer set (two steps old)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            for h, o in zip(out_host[slot], outs):
                h.copy_(o, non_blocking=True)
                o.record_stream(copy_stream)
            copied[slot].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, tail=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if tail is not None:
            tail()              # e.g. make the timing stream wait for side-stream copies of the last step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        device_step(pts_dev)
        e2e_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(lambda: device_step(pts_dev), args.steps)
    ms_e2e = timed(e2e_step, args.steps, tail=lambda: torch.cuda.current_stream().wait_stream(copy_stream))

    # forward only (graph + features already built): the GNN proper, for edges/s
    with torch.no_grad():
        ms_fwd = timed(lambda: det.forward_batch(bf0.gb, bf0.node_features, bf0.edge_features, training=False), args.steps)
    ms_graph = timed(lambda: gf.build_graph_batch(pts_dev, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH), args.steps)
    roof = measure_roofline(det, bf0, dev, args.steps)
    train = measure_train(dev, args, timed, rank, world) if not args.no_train else None
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf, which = peaks()
    total_frames = args.frames * world
    per_step_s = ms_dev / args.steps / 1e3
    line = {
        'metric': 'radar_frames_per_s_gnn_fwd', 'value': total_frames / per_step_s, 'unit': 'frames/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args),
        'edges_per_s_gnn_fwd': n_edges * world / (ms_fwd / args.steps / 1e3),
        'frames_per_s_gnn_fwd_only': total_frames / (ms_fwd / args.steps / 1e3),
        'breakdown_ms': {'graph_build_and_features': ms_graph / args.steps, 'gnn_forward': ms_fwd / args.steps},
        'graph': {'nodes_per_gpu': n_nodes, 'directed_edges_per_gpu': n_edges, 'undirected_links_per_gpu': n_und,
                  'clusters_per_gpu': n_clusters},
        'e2e': {'value': total_frames / (ms_e2e / args.steps / 1e3), 'unit': 'frames/s',
                'h2d_bytes_per_step': int(sum(v.numel() * v.element_size() for v in host.values())),
                'd2h_bytes_per_step': int(sum(h.numel() * h.element_size() for h in out_host[0])),
                'overlap': 'read-back of step i on a side stream under the compute of step i+1 (double-buffered pinned results)'},
        'gpu_launches': launches_per_step(det),
        'roofline': dict(roof, peak=hbm, frac=roof['achieved'] / hbm, peak_source=which),
        'clocks': clocks,
    }
    if train:
        line['train'] = train
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def launches_per_step(det):
    # kernels only (memsets excluded).  graph build: knn, sym_count, add, 3 scan, copy_last, sym_fill, sort, finalize
    # (rows, 3 scan, und, copy_last), node feat, edge feat = 17; weight packing: 2 pack_kernel + 2 pack_split per conv
    # block; model: node enc, edge enc, per conv block (tcgen05 edge kernel, node program), 6 head programs
    L = len(det.pass_messages.conv_blk)
    return 17 + 2 + 2 * L + 2 + 2 * L + 6


def _traffic_from_profile(n_edges, n_nodes):
    """DRAM bytes per launch of mp_edge_tc_kernel from the committed `ncu --set full` capture (profiles/roofline_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum at a stated graph size), scaled linearly to this run's graph."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if not os.path.exists(path):
        return None
    t = json.load(open(path))
    alg_here = 260.0 * n_edges + 1280.0 * n_nodes
    alg_there = 260.0 * t['n_edges'] + 1280.0 * t['n_nodes']
    return t['dram_bytes'] * alg_here / alg_there


def measure_roofline(det, bf, dev, steps):
    """Dominant kernel of the forward: mp_edge_tc_kernel (message function + aggregation of one conv block), timed alone on
    its stream with CUDA events.  Algorithmic bytes of that kernel: 260 E (edge embedding row + target/source index) +
    1280 N (hoisted projection rows read once, aggregated messages written once).  The whole layer (projection, edge kernel,
    node update; SURVEY.md 8d: B_f = 512 N + 260 E) is reported beside it."""
    import ctypes as C
    from graph_neural_network_for_radar_perception_b200._engine import detector_table
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr
    gb = bf.gb
    table = detector_table(det)
    table.refill(None)
    check(lib().rgnn_pack_detector(C.byref(table.det), stream_ptr()), 'pack')
    N, E = gb.n_nodes, gb.n_edges
    x = torch.randn(N, 64, device=dev)
    e = torch.randn(E, 64, device=dev)
    out = torch.empty_like(x)
    agg = torch.empty_like(x)
    proj = torch.empty(N, 256, device=dev)
    g = gb.c_struct()
    conv = table.det.conv[0]
    s = stream_ptr()
    reps = max(steps, 3)

    def timed_ms(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_layer = timed_ms(lambda: check(lib().rgnn_conv_block_fwd(C.byref(conv), C.byref(g), ptr(x), ptr(e), ptr(out), ptr(agg), ptr(proj), s), 'conv'))
    # the edge kernel alone (its launch also zero-fills agg with a memset node, counted in its time)
    ms_edge = timed_ms(lambda: check(lib().rgnn_conv_edges_fwd(C.byref(conv), C.byref(g), ptr(e), ptr(proj), ptr(agg), s), 'edges'))
    bytes_edge = 260.0 * E + 1280.0 * N
    bytes_layer = 512.0 * N + 260.0 * E
    flops_alg = 65536.0 * E + 16384.0 * N
    return {'kernel': 'mp_edge_tc_kernel (tcgen05 3xTF32: message function + segmented-sum aggregation of one conv block)',
            'bound': 'hbm',
            'achieved': bytes_edge / (ms_edge * 1e-3) / 1e9, 'unit': 'GB/s', 'traffic': _traffic_from_profile(E, N),
            'ms_per_launch': ms_edge, 'algorithmic_bytes': bytes_edge,
            'algorithmic_tflops': 65536.0 * E / (ms_edge * 1e-3) / 1e12,
            'executed_tf32_tflops': 3 * 2 * 16384.0 * E / (ms_edge * 1e-3) / 1e12,
            'note': 'fp32-parity mode executes 3 TF32 MMAs per product and is bound by the tensor pipe + CUDA-core epilogue, not by HBM',
            'layer': {'ms': ms_layer, 'algorithmic_bytes': bytes_layer, 'achieved_GBps': bytes_layer / (ms_layer * 1e-3) / 1e9,
                      'algorithmic_tflops': flops_alg / (ms_layer * 1e-3) / 1e12}}


def measure_train(dev, args, timed, rank, world):
    """BASELINE.json configs[2]: full multi-task training step (forward, 4 losses, backward, gradient all-reduce
    over NCCL when world > 1, fused SGD) on `--train-frames` frames per GPU (weak scaling)."""
    from graph_neural_network_for_radar_perception_b200 import config, Model_Training
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer
    model = Model_Training(config(), dev)
    model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True))
    model = model.to(dev).train()
    trainer = DataParallelTrainer(model)
    nfr = args.train_frames
    frames = make_frames(nfr, args.points, seed0=5000 + 1000 * rank)
    pts, fp = gf.frames_to_device([f[0] for f in frames], dev)
    bf = gf.build_graph_batch(pts, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH)
    gb = bf.gb
    # synthetic labels (synth.make_labels needs the edge list of each frame: take it from the GPU graph)
    ei = bf.edge_index().cpu().numpy()
    row_ptr = gb.row_ptr.cpu().numpy()
    labs, cl_lists = [], []
    for i, (d, src) in enumerate(frames):
        e0, e1 = int(row_ptr[fp[i]]), int(row_ptr[fp[i + 1]])
        lab = synth.make_labels(d, src, ei[:, e0:e1] - fp[i])
        labs.append(lab)
        cl_lists.append([torch.from_numpy(c) for c in lab['cluster_node_idx']])
    gb.set_clusters(cl_lists, fp[:-1], dev)
    labels = {k: torch.cat([torch.from_numpy(l[k]) for l in labs]).to(dev)
              for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
    step = lambda: trainer.step(gb, bf.node_features, bf.edge_features, labels)
    for _ in range(max(args.warmup, 3)):
        step()
    ms = timed(step, args.steps) / args.steps
    loss, _ = step()
    return {'metric': 'radar_frames_per_s_gnn_fwd_bwd', 'value': nfr * world / (ms * 1e-3), 'unit': 'frames/s',
            'ms_per_step': ms, 'frames_per_gpu': nfr, 'edges_per_s': gb.n_edges * world / (ms * 1e-3),
            'directed_edges_per_gpu': gb.n_edges, 'collective': 'nccl all-reduce of 463144 fp32 gradients + 3 counts' if world > 1 else 'none (1 GPU)',
            'includes': 'forward + 4 losses + backward + SGD(momentum, weight decay) update; graph prebuilt',
            'loss_after': float(sum(v.item() for v in loss.values()))}


def cpu_baseline(args):
    """The reference's CPU path (oracle port, pinned to the reference's own outputs) on this box's host cores, on a bounded
    sample of the same workload: graph construction, eval forward and a training step (forward + losses + backward) timed
    separately, plus the reference's own CPU-runnable case (BASELINE.json configs[0]: one frame of ~1000 points)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = torch.load(CKPT, map_location='cpu', weights_only=True)
    sample = max(1, min(args.ref_frames, args.frames))
    frames = make_frames(sample, args.points)
    cpu_frame_pass(sd, *frames[0])
    times = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    edges = reps = 0
    while True:
        for f in frames:
            edges += cpu_frame_pass(sd, *f, times=times)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 3:
            break
    dt = time.perf_counter() - t0
    n = reps * sample
    # training step, per-frame loop as Model_Training.forward does it (2-frame sample)
    tt = [0.0, 0.0, 0.0]
    tr_frames = frames[:2]
    t1 = time.perf_counter()
    tr_edges = sum(cpu_frame_pass(sd, *f, train=True, times=tt) for f in tr_frames)
    dtr = time.perf_counter() - t1
    # configs[0]: one frame of 1000 points
    f1k = make_frames(1, 1000)[0]
    cpu_frame_pass(sd, *f1k)
    t1k = [0.0, 0.0, 0.0]
    for _ in range(3):
        cpu_frame_pass(sd, *f1k, times=t1k)
    return {'value': n / dt, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'edges_per_s': edges / dt,
            'sample': f'{reps} x {sample} frames x {args.points} points (graph build + forward per frame, oracle port of the reference)',
            'ms_per_frame': {'graph_build_and_features': 1e3 * times[0] / n, 'forward': 1e3 * times[1] / n},
            'train': {'value': len(tr_frames) / dtr, 'unit': 'frames/s', 'edges_per_s': tr_edges / dtr,
                      'sample': f'{len(tr_frames)} frames x {args.points} points, forward + 4 losses + backward per frame (no optimizer step)',
                      'ms_per_frame': {'graph_build_and_features': 1e3 * tt[0] / len(tr_frames), 'forward_and_loss': 1e3 * tt[1] / len(tr_frames),
                                       'backward': 1e3 * tt[2] / len(tr_frames)}},
            'c1_one_frame_1000_points_ms': {'graph_build_and_features': 1e3 * t1k[0] / 3, 'forward': 1e3 * t1k[1] / 3}}


_JSON_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE line, the JSON.  Libraries write there too (NCCL prints its version banner to stdout
    from C), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the original descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--frames', type=int, default=256)
    ap.add_argument('--points', type=int, default=3000)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--ref-frames', type=int, default=4)
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--train-frames', type=int, default=64)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
