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
    # inputs: two device buffer sets; the H2D copy of step i + 1 runs on its own stream under the compute of step i (one full
    # copy per step, like the read-back: a serving loop uploads the next request while the current one computes)
    in_stream = torch.cuda.Stream(device=dev)
    in_dev = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    in_ready = [torch.cuda.Event(), torch.cuda.Event()]
    in_free = [None, None]
    in_pending = [False, False]
    e2e_count = 0

    def upload(slot):
        with torch.cuda.stream(in_stream):
            if in_free[slot] is not None:
                in_stream.wait_event(in_free[slot])     # the step that read this buffer set has finished with it
            for k, v in host.items():
                in_dev[slot][k].copy_(v, non_blocking=True)
            in_ready[slot].record()
        in_pending[slot] = True

    def e2e_step():
        nonlocal e2e_count
        slot = e2e_count & 1
        e2e_count += 1
        if not in_pending[slot]:
            upload(slot)                                # first call only: nothing was uploaded ahead
        torch.cuda.current_stream().wait_event(in_ready[slot])
        in_pending[slot] = False
        upload(slot ^ 1)                                # the NEXT step's inputs, under this step's compute
        outs = device_step(in_dev[slot])
        in_free[slot] = torch.cuda.Event()
        in_free[slot].record()
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

    def timed_local(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

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
    ms_e2e = timed(e2e_step, args.steps, tail=lambda: (torch.cuda.current_stream().wait_stream(copy_stream), torch.cuda.current_stream().wait_stream(in_stream)))

    # forward only (graph + features already built): the GNN proper, for edges/s
    with torch.no_grad():
        ms_fwd = timed(lambda: det.forward_batch(bf0.gb, bf0.node_features, bf0.edge_features, training=False), args.steps)
    ms_graph = timed(lambda: gf.build_graph_batch(pts_dev, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH), args.steps)
    hbm, tf, which = peaks()
    roof, roof_extra = measure_roofline(det, bf0, dev, args.steps, hbm, tf)
    # reduced-precision mode (single fp16 pass, fp32 accumulate): its own line, never the headline
    reduced = None
    if rank == 0:
        with torch.no_grad():
            ref_out = det.forward_batch(bf0.gb, bf0.node_features, bf0.edge_features, training=False)
            _cabi.check(_cabi.lib().rgnn_set_option(b'f16_passes', 1), 'opt')
            try:
                ms_fwd_1p = timed_local(lambda: det.forward_batch(bf0.gb, bf0.node_features, bf0.edge_features, training=False), args.steps)
                out_1p = det.forward_batch(bf0.gb, bf0.node_features, bf0.edge_features, training=False)
            finally:
                _cabi.check(_cabi.lib().rgnn_set_option(b'f16_passes', 3), 'opt')
        rel = [float(((a - b).abs().max() / b.abs().max().clamp_min(1e-9)).item()) for a, b in zip(out_1p, ref_out)]
        reduced = {'dtype': 'f16 operands (single pass), f32 accumulate', 'gnn_forward_ms': ms_fwd_1p / args.steps,
                   'frames_per_s_gnn_fwd_only': args.frames / (ms_fwd_1p / args.steps / 1e3),
                   'max_abs_err_over_max_abs_vs_fp32_path': dict(zip(('node_cls', 'node_off', 'link_cls', 'obj_cls'), rel)),
                   'stated_tolerance': '5e-3 of each output tensor\'s largest magnitude (tests/test_model_gpu.py::test_reduced_precision_mode)',
                   'scope': 'edge encoder, 7 message kernels, stems / heads run single-pass; node encoder, node update and the per-cluster head stay 3xTF32'}
    # SURVEY 8(d) C2, second graph: the radius-union adjacency of compute_adjacency_information_v2 (graph_features.py:87-114) on the
    # same frames -- about twice the edges; graph build + forward like the headline step (rank 0, its own GPU)
    union = None
    if rank == 0:
        def union_step():
            bfu = gf.build_graph_batch(pts_dev, fp, EPS2, KNN, union_radius=True, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH)
            bfu.gb.cl_ptr, bfu.gb.cl_members, bfu.gb.n_clusters = cl_ptr, cl_members, n_clusters
            with torch.no_grad():
                det.forward_batch(bfu.gb, bfu.node_features, bfu.edge_features, training=False)
            return bfu.gb.n_edges
        e_union = union_step()
        ms_union = timed_local(union_step, max(2, min(args.steps, 3)))
        union = {'workload': 'the same 256 frames with the radius-union graph (kNN k=10 OR d^2 <= 25: compute_adjacency_information_v2)',
                 'directed_edges_per_gpu': e_union, 'ms_per_step': ms_union / max(2, min(args.steps, 3)),
                 'value': args.frames / (ms_union / max(2, min(args.steps, 3)) / 1e3), 'unit': 'frames/s (this GPU)',
                 'edges_per_s': e_union / (ms_union / max(2, min(args.steps, 3)) / 1e3)}
    c1 = measure_c1(det, model, dev, train_step=(world == 1)) if rank == 0 else None
    parity = parity_check(det, bf0, frames, fp, cl_lists, ref_out) if rank == 0 else None
    train = measure_train(dev, args, timed, rank, world) if not args.no_train else None
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
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
                'overlap': 'upload of step i+1 and read-back of step i-1 on side streams under the compute of step i (double-buffered device inputs and pinned results; one full H2D and one full D2H per step inside the timed region)'},
        'gpu_launches': launches_per_step(det),
        'roofline': dict(roof, peak_source=which),
        'clocks': clocks,
    }
    line.update({k: dict(v, peak_source=which) for k, v in roof_extra.items()})
    line['reduced_precision'] = reduced
    line['c1_latency'] = c1
    line['radius_union_graph'] = union
    line['parity_max_err'] = parity[0]
    line['parity'] = {'what': 'frame 0 of the bench batch, four outputs vs the float32 oracle; ratio to (1e-4 |want| + 1e-5 max|want|), <= 1 passes',
                      'max_abs_err': parity[1]}
    if train:
        line['train'] = train
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def launches_per_step(det):
    # kernels only (memsets excluded), one device_step.  graph build: knn, sym_count, add, 3 scan, copy_last, sym_fill, sort,
    # finalize (rows, 3 scan, und, copy_last), node feat, edge feat = 17; weight images are cached between steps (0);
    # model: node encoder, edge encoder, per conv block (mp_edge_f16_kernel, node program), 6 stem / head kernels
    L = len(det.pass_messages.conv_blk)
    return 17 + 2 + 2 * L + 6


def _traffic_from_profile(n_edges, n_nodes):
    """DRAM bytes per launch of the message kernel from the committed `ncu --set full` capture (profiles/roofline_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum at a stated graph size), scaled linearly to this run's graph."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if not os.path.exists(path):
        return None
    t = json.load(open(path))
    alg_here = 260.0 * n_edges + 512.0 * n_nodes
    alg_there = 260.0 * t['n_edges'] + 512.0 * t['n_nodes']
    return t['dram_bytes'] * alg_here / alg_there


def _traffic_entry(key, n_edges, n_nodes, per_edge, per_node):
    """Same for the other captured kernels (`edge_enc`: the inference edge encoder; `bwd`: mp_edge_bwd_f16_kernel + dproj_gather_kernel of one
    conv block), scaled by the algorithmic bytes per_edge * E + per_node * N."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if not os.path.exists(path):
        return None
    t = json.load(open(path)).get(key)
    if not t:
        return None
    return t['dram_bytes'] * (per_edge * n_edges + per_node * n_nodes) / (per_edge * t['n_edges'] + per_node * t.get('n_nodes', 0))


def measure_roofline(det, bf, dev, steps, hbm, tf):
    """Roofline entries with SURVEY.md section 8(d)'s ALGORITHMIC figures (fp32 storage, int32 indices, weights counted once):
      per conv layer forward   B_f = 512 N + 260 E bytes,  F_f = 65 536 E + 16 384 N FLOP   (message path alone: 65 536 E)
      per conv layer backward  B_b = 768 N + 772 E bytes,  2 F_f FLOP (the recompute is not counted)
      edge encoder             2 x 59 136 FLOP and 28 B in + 256 B out per edge
    Every kernel is timed alone on its stream with CUDA events, through the C-ABI entry the detector itself calls.
    In the fp32-parity mode the message kernel executes three fp16 MMAs per product, which puts the tensor pipe (0.67 ms for
    9.4 M edges at the measured sustained peak) above the HBM floor (0.43 ms): `bound` says so; the HBM view is reported
    beside it, and the single-pass mode (HBM-bound) has its own entry."""
    import ctypes as C
    from graph_neural_network_for_radar_perception_b200._engine import detector_table
    from graph_neural_network_for_radar_perception_b200._cabi import check, lib, ptr, stream_ptr
    gb = bf.gb
    table = detector_table(det)
    table.refill(None)
    table.ensure_packed(stream_ptr())
    N, E = gb.n_nodes, gb.n_edges
    x = torch.randn(N, 64, device=dev)
    e = torch.randn(E, 64, device=dev)
    es = torch.empty(lib().rgnn_split_edge_embedding_words(E), dtype=torch.int32, device=dev)
    check(lib().rgnn_split_edge_embedding(ptr(e), E, ptr(es), stream_ptr()), 'split')
    out, agg = torch.empty_like(x), torch.empty_like(x)
    proj, proj2 = torch.randn(N, 256, device=dev), torch.empty(N, 256, device=dev)
    g = gb.c_struct()
    conv, conv_next = table.det.conv[0], table.det.conv[1]
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

    L = lib()
    edges = lambda: check(L.rgnn_conv_edges_f16_fwd(C.byref(conv), C.byref(g), ptr(es), ptr(proj), ptr(agg), s), 'edges')
    layer = lambda: check(L.rgnn_conv_layer_f16_fwd(C.byref(conv), C.byref(conv_next), C.byref(g), ptr(x), ptr(es), ptr(proj), ptr(out),
                                                    ptr(agg), ptr(proj2), s), 'layer')
    ms_edge = timed_ms(edges)        # its launch also zero-fills agg with a memset node, counted in its time
    ms_layer = timed_ms(layer)
    check(L.rgnn_set_option(b'f16_passes', 1), 'opt')
    try:
        ms_edge_1p = timed_ms(edges)
        ms_layer_1p = timed_ms(layer)
    finally:
        check(L.rgnn_set_option(b'f16_passes', 3), 'opt')
    # edge encoder
    ef = bf.edge_features
    ms_enc = timed_ms(lambda: check(L.rgnn_edge_encoder_f16_fwd(C.byref(table.det.edge_enc), ptr(ef), ptr(gb.perm), E, ptr(es), None, s), 'enc'))
    # message backward of one layer (dgrad kernel + the two weight-gradient GEMMs + the projection-gradient gather)
    nb = L.rgnn_conv_msg_bwd_workspace_bytes(C.byref(conv), C.byref(g))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    dagg, dproj, de = torch.randn(N, 64, device=dev) * 1e-3, torch.empty(N, 256, device=dev), torch.empty(E, 64, device=dev)
    ms_bwd = timed_ms(lambda: check(L.rgnn_conv_msg_f16_bwd(C.byref(conv), C.byref(g), ptr(es), ptr(proj), ptr(dagg), ptr(dproj), ptr(de),
                                                            ptr(ws), nb, s), 'bwd'))
    B_f, B_b = 512.0 * N + 260.0 * E, 768.0 * N + 772.0 * E
    F_msg, F_f = 65536.0 * E, 65536.0 * E + 16384.0 * N
    gbs = lambda b, ms: b / (ms * 1e-3) / 1e9
    tfs = lambda f, ms: f / (ms * 1e-3) / 1e12
    main = {'kernel': 'mp_edge_f16_kernel (tcgen05 kind::f16, fp16-split operands x 3 passes = fp32 parity: message function + '
                      'segmented-sum aggregation of one conv block)',
            'bound': 'tensor', 'achieved': tfs(F_msg, ms_edge), 'peak': tf, 'unit': 'TFLOP/s', 'frac': tfs(F_msg, ms_edge) / tf,
            'traffic': _traffic_from_profile(E, N), 'ms_per_launch': ms_edge,
            'algorithmic_flop': F_msg, 'executed_f16_tflops': tfs(3 * 2 * 16384.0 * E, ms_edge),
            'hbm_view': {'algorithmic_bytes': B_f, 'kernel_GBps': gbs(B_f, ms_edge), 'kernel_frac_of_hbm': gbs(B_f, ms_edge) / hbm,
                         'layer_ms': ms_layer, 'layer_GBps': gbs(B_f, ms_layer), 'layer_frac_of_hbm': gbs(B_f, ms_layer) / hbm,
                         'layer_algorithmic_tflops': tfs(F_f, ms_layer)},
            'note': 'three fp16 MMAs per product (fp32 parity) put the tensor floor above the HBM floor; the single-pass entry is HBM-bound'}
    extra = {
        'roofline_reduced': {'kernel': 'mp_edge_f16_kernel, single pass (fp16 operands, fp32 accumulate; rgnn_set_option f16_passes=1)',
                             'bound': 'hbm', 'achieved': gbs(B_f, ms_edge_1p), 'peak': hbm, 'unit': 'GB/s', 'frac': gbs(B_f, ms_edge_1p) / hbm,
                             'ms_per_launch': ms_edge_1p, 'layer_ms': ms_layer_1p, 'layer_frac': gbs(B_f, ms_layer_1p) / hbm, 'traffic': None},
        'roofline_rowmlp': {'kernel': 'edge_enc_f16_kernel (graph_feature_encoding of the edges, 7-256-128-128-64)', 'bound': 'tensor',
                            'achieved': tfs(2 * 59136.0 * E, ms_enc), 'peak': tf, 'unit': 'TFLOP/s', 'frac': tfs(2 * 59136.0 * E, ms_enc) / tf,
                            'ms_per_launch': ms_enc, 'algorithmic_bytes': 284.0 * E, 'traffic': _traffic_entry('edge_enc', E, N, 284.0, 0.0)},
        'roofline_bwd': {'kernel': 'message backward of one conv block: absmax_kernel + mp_edge_bwd_f16_kernel (recompute + dgrad + both weight gradients fused, '
                                   'fp16-split x 3 passes) + dproj_gather_kernel',
                         'bound': 'tensor', 'achieved': tfs(2 * F_msg, ms_bwd), 'peak': tf, 'unit': 'TFLOP/s', 'frac': tfs(2 * F_msg, ms_bwd) / tf,
                         'ms': ms_bwd, 'hbm_view': {'algorithmic_bytes': B_b, 'GBps': gbs(B_b, ms_bwd), 'frac_of_hbm': gbs(B_b, ms_bwd) / hbm},
                         'traffic': _traffic_entry('bwd', E, N, 772.0, 768.0),
                         'traffic_note': 'the 512 B / edge of dz1 written by the fused kernel and read twice by dproj_gather_kernel put the traffic '
                                         'at ~2.3 x the algorithmic bytes'},
    }
    return main, extra


def measure_c1(det, model, dev, train_step=True):
    """The reference's own calling convention (one frame at a time, modules/inference/output.py:88-94; lists of small frames
    in Model_Training.forward): host + device latency in ms, median of 20 calls, everything a caller pays included
    (int64 edge_index -> CSR, cluster lists, weight images cached).  C1 = BASELINE.json configs[0]'s frame shape."""
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf
    out = {}

    def lat(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        return float(np.median(ts))
    # 1 frame x 1000 points through Model_Inference.forward(node_features, edge_features, edge_index, adj, clusters)
    d, src = synth.make_frame(4242, 1000, knn=KNN)
    pts, fp = gf.frames_to_device([d], dev)
    bf = gf.build_graph_batch(pts, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH)
    ei = bf.edge_index()
    blob = src['blob_of']
    cl = [torch.from_numpy(np.nonzero(blob == b)[0]).to(dev) for b in np.unique(blob[blob >= 0])] + \
         [torch.tensor([i], device=dev) for i in np.nonzero(blob < 0)[0]]
    with torch.no_grad():
        out['inference_1_frame_1000_points_ms'] = lat(lambda: det(bf.node_features, bf.edge_features, ei, None, cl))
        out['graph_build_1_frame_1000_points_ms'] = lat(lambda: gf.build_graph_batch(pts, fp, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH))
        # the reference's inference call proper (inference/output.py:88-94: no cluster lists; offsets -> centres -> connected components of
        # the predicted links on the device -> class head on the found clusters), member lists returned as the reference returns them
        other = torch.from_numpy(np.stack([d['meas_px'], d['meas_py'], d['meas_vx'], d['meas_vy']], axis=1).astype(np.float32)).to(dev)
        had = getattr(det, 'extract_proposals', False)
        det.set_param_for_proposal_extraction(1.4, True)
        out['inference_with_proposal_extraction_1_frame_1000_points_ms'] = lat(lambda: det(bf.node_features, bf.edge_features, ei, None, None, other))
        det.extract_proposals = had
    # 8 frames x 100 points through Model_Training.forward with per-frame lists (forward + losses; eval mode, no backward)
    nf_l, ef_l, ei_l = [], [], []
    labels = {k: [] for k in ('cluster_node_idx', 'cluster_labels', 'edge_class', 'node_class', 'node_offsets')}
    for i in range(8):
        d, src = synth.make_frame(4300 + i, 100, knn=KNN)
        p1, f1 = gf.frames_to_device([d], dev)
        b1 = gf.build_graph_batch(p1, f1, EPS2, KNN, max_range=GRID_MAX_R, max_azimuth=GRID_MAX_TH)
        e1 = b1.edge_index()
        lab = synth.make_labels(d, src, e1.cpu().numpy())
        nf_l.append(b1.node_features); ef_l.append(b1.edge_features); ei_l.append(e1)
        labels['cluster_node_idx'].append([torch.from_numpy(c).to(dev) for c in lab['cluster_node_idx']])
        for k in ('cluster_labels', 'edge_class', 'node_class', 'node_offsets'):
            labels[k].append(torch.from_numpy(lab[k]).to(dev))
    with torch.no_grad():
        out['training_forward_8_frames_100_points_ms'] = lat(lambda: model(nf_l, ef_l, ei_l, [None] * 8, labels))
    if not train_step:      # rank 0 alone is here: under torchrun the trainer's all-reduce would wait for ranks that never come (1-GPU figure only)
        return out
    # the same 8 small frames through a whole optimisation step (pack_batch + DataParallelTrainer.step: forward, losses, backward,
    # SGD; a fresh copy of the model, its parameters are re-homed in the trainer's flat buffers); steps are issued back to back
    from graph_neural_network_for_radar_perception_b200 import config as _config, Model_Training as _MT
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer
    m2 = _MT(_config(), dev)
    m2.load_state_dict(model.state_dict())
    m2 = m2.to(dev).train()
    trainer = DataParallelTrainer(m2, lr=1e-5)

    def train_step():
        gb, x, e = m2.pack_batch(nf_l, ef_l, ei_l, labels['cluster_node_idx'])
        trainer.step(gb, x, e, labels)
    for _ in range(5):
        train_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        train_step()
    torch.cuda.synchronize()
    out['training_step_8_frames_100_points_ms'] = 1e3 * (time.perf_counter() - t0) / 50
    return out


def parity_check(det, bf, frames, fp, cl_lists, outs):
    """CHECKER (outside every timed region; the one place the product arm touches oracle/): the four outputs of frame 0 of the
    bench batch against the float32 oracle of the reference.  Returns the largest |got - want| / (1e-4 |want| + 1e-5 max|want|)
    (<= 1 means inside the tolerance of tests/test_model_gpu.py) and the raw maxima."""
    from oracle import graph_np, model_torch
    sd = torch.load(CKPT, map_location='cpu', weights_only=True)
    d = frames[0][0]
    adj = graph_np.adjacency_information(d, EPS2, KNN)
    nf = torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, GRID_MAX_R, 0, GRID_MAX_TH).astype(np.float32))
    ef = torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32))
    with torch.no_grad():
        want = model_torch.detector_forward(sd, nf, ef, torch.from_numpy(adj['adj_list']), [c for c in cl_lists[0]])
    n0, eu0, c0 = fp[1], adj['adj_list'].shape[1] // 2, len(cl_lists[0])
    got = (outs[0][:n0], outs[1][:n0], outs[2][:eu0], outs[3][:c0])
    worst, raw = 0.0, {}
    for name, gt, w in zip(('node_cls', 'node_off', 'link_cls', 'obj_cls'), got, want):
        w = w.numpy().astype(np.float64)
        err = np.abs(gt.cpu().numpy().astype(np.float64) - w)
        tol = 1e-4 * np.abs(w) + max(1e-5 * np.abs(w).max(), 2e-6)
        worst = max(worst, float((err / tol).max()))
        raw[name] = float(err.max())
    return worst, raw


def measure_train(dev, args, timed, rank, world):
    """BASELINE.json configs[2] / [3]: full multi-task training step (forward, 4 losses, backward, ONE gradient all-reduce over
    NCCL when world > 1 -- the counts are all-reduced on the stream, nothing is read back by the host -- fused SGD).
    weak: `--train-frames` frames per GPU; strong (world > 1): the SAME global batch of `--train-frames` frames split over
    the ranks."""
    from graph_neural_network_for_radar_perception_b200 import config, Model_Training
    from graph_neural_network_for_radar_perception_b200 import graph_features as gf
    from graph_neural_network_for_radar_perception_b200.training import DataParallelTrainer, shard_range

    def case(frames):
        model = Model_Training(config(), dev)
        model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=True))
        model = model.to(dev).train()
        trainer = DataParallelTrainer(model)
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
        return ms, gb.n_edges, float(sum(v.item() for v in loss.values()))

    nfr = args.train_frames
    ms, n_edges, loss = case(make_frames(nfr, args.points, seed0=5000 + 1000 * rank))
    out = {'metric': 'radar_frames_per_s_gnn_fwd_bwd', 'value': nfr * world / (ms * 1e-3), 'unit': 'frames/s',
           'ms_per_step': ms, 'frames_per_gpu': nfr, 'edges_per_s': n_edges * world / (ms * 1e-3),
           'directed_edges_per_gpu': n_edges, 'scaling': 'weak',
           'collective': 'nccl: all-reduce of 3 counts (on the stream) + ONE all-reduce of 463144 fp32 gradients with the NaN flag '
                         'and the four loss shares behind them' if world > 1 else 'none (1 GPU)',
           'includes': 'forward + 4 losses + backward + SGD(momentum, weight decay) update; graph prebuilt; no host read in the step',
           'loss_after_global_batch': loss}
    if world > 1:
        allf = make_frames(nfr, args.points, seed0=5000)
        mine = [allf[i] for i in shard_range(nfr, rank, world)]
        ms_s, _, loss_s = case(mine)
        out['strong'] = {'scaling': 'strong', 'global_frames': nfr, 'frames_per_gpu': len(mine), 'ms_per_step': ms_s,
                         'value': nfr / (ms_s * 1e-3), 'unit': 'frames/s', 'loss_after_global_batch': loss_s}
    return out


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
