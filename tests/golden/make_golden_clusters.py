"""Golden fixtures for the proposal clustering: runs the UNMODIFIED reference
modules/inference/clustering.py::Simple_DBSCAN (NumPy only) from /root/reference in the build container.

    python tests/golden/make_golden_clusters.py      ->  tests/golden/clusters.npz

Cases: predicted centres = noisy blob centres (so that clusters of several nodes exist), the undirected pair list of a
symmetrised kNN graph built by the reference's own compute_adjacency_information, random 0/1 link predictions biased
towards "same blob"; both modes of the class (links with the distance gate eps, offsets with the squared-distance eps).
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    clus = load(os.path.join(REF, 'modules/inference/clustering.py'), 'ref_clustering')
    gfeat = load(os.path.join(REF, 'modules/compute_features/graph_features.py'), 'ref_graph_features')
    sys.path.insert(0, REF)
    import modules.inference.inference as infer          # the reference package itself (NumPy / torch only)
    from graph_neural_network_for_radar_perception_b200 import synth
    out = {}
    for case, (seed, n, eps_l, eps_r) in enumerate([(11, 60, 1.4, 1.4), (12, 300, 2.0, 0.6), (13, 7, 1.4, 1.4), (14, 500, 0.8, 3.0)]):
        d, src = synth.make_frame(seed, n)
        rng = np.random.default_rng(seed)
        adj = gfeat.compute_adjacency_information(d, 25, 10)
        blob = src['blob_of']
        centres = np.stack([d['meas_px'], d['meas_py']], axis=1).astype(np.float32)
        for b in np.unique(blob[blob >= 0]):
            centres[blob == b] = centres[blob == b].mean(axis=0) + rng.normal(0, 0.3, size=(int((blob == b).sum()), 2)).astype(np.float32)
        r, c = np.nonzero(np.triu(adj['adj_matrix'], k=1))
        same = (blob[r] == blob[c]) & (blob[r] >= 0)
        pred = np.where(rng.random(r.shape[0]) < 0.9, same, ~same).astype(np.int64)
        o = clus.Simple_DBSCAN(eps_l, True)
        o.cluster_nodes(centres.copy(), pred.copy(), adj['adj_matrix'].copy())
        ids_l, n_l = o.meas_to_cluster_id.astype(np.int64).copy(), o.num_clusters
        o = clus.Simple_DBSCAN(eps_r, False)
        o.cluster_nodes(centres.copy())
        ids_r, n_r = o.meas_to_cluster_id.astype(np.int64).copy(), o.num_clusters
        p = f'c{case}_'
        out.update({p + 'centres': centres, p + 'und': np.stack([r, c]).astype(np.int64), p + 'pred': pred,
                    p + 'adj_matrix': adj['adj_matrix'], p + 'eps_links': np.float64(eps_l), p + 'eps_radius': np.float64(eps_r),
                    p + 'ids_links': ids_l, p + 'n_links': np.int64(n_l), p + 'ids_radius': ids_r, p + 'n_radius': np.int64(n_r)})
        print(f'case {case}: n={n} pairs={r.shape[0]} clusters links={n_l} radius={n_r}')
        # proposals of the links-mode clusters: the reference's compute_proposals (modules/inference/inference.py:36-47) and
        # the majority vote of output.py:111-118 over random segmentation logits
        import torch
        members = [torch.from_numpy(np.nonzero(ids_l == i)[0]) for i in range(n_l)]
        noise = 0.5 * np.eye(2, dtype=np.float32)
        mu, sig, size = infer.compute_proposals(members, centres[:, 0].copy(), centres[:, 1].copy(), noise)
        node_logits = rng.normal(size=(n, 7)).astype(np.float32)
        pred = torch.from_numpy(node_logits).argmax(dim=-1)
        vote = np.array([int(torch.argmax(torch.bincount(pred[m]))) for m in members], dtype=np.int64)
        out.update({p + 'prop_mean': np.stack(mu).astype(np.float32), p + 'prop_cov': np.stack(sig).astype(np.float32),
                    p + 'prop_size': np.array(size, dtype=np.int64), p + 'node_logits': node_logits, p + 'prop_vote': vote,
                    p + 'prop_dtypes': np.array([str(np.stack(mu).dtype), str(np.stack(sig).dtype)])})
    out['n_cases'] = np.int64(4)
    np.savez_compressed(os.path.join(HERE, 'clusters.npz'), **out)


if __name__ == '__main__':
    main()
