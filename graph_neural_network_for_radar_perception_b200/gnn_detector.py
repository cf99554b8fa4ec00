"""`Model_Inference`, `Model_Training`, `Model_Object_Classifier_Finetuning` with the reference's
constructor / forward signatures and state_dict layout (reference modules/neural_net/gnn/gnn_detector.py),
running on the fused CUDA tile programs of librgnn.so.

Differences a caller can observe (all documented in DESIGN.md):
  * tensors and parameters must be on a CUDA device (no CPU path);
  * `adj_matrix` is accepted for signature compatibility but never read: the undirected link list is
    `edge_index[:, src < dst]`, identical to nonzero(triu(adj_matrix, 1)) for reference-built graphs;
  * `Model_Training.forward` packs all frames of the batch into one block-diagonal graph and runs the
    detector once (bit-identical per frame because channel_normalization is strictly per row) instead of the
    reference's Python loop over frames (gnn_detector.py:443-452).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from ._engine import GraphBatch, run_detector, run_detector_then_cluster
from .clustering import Simple_DBSCAN
from .compute_offsets import normalize_gt_offsets, unnormalize_gt_offsets
from .gnn_blocks import (graph_convolution, graph_feature_encoding, link_predictions, node_offset_predictions,
                         node_predictions, node_segmentation, object_classification)
from .loss import Loss_Graph, Loss_Object_Class


@torch.no_grad()
def compute_accuracy(predicted_class, gt_class):
    """Arg-max hit rate (reference gnn_detector.py:23-28)."""
    return (predicted_class.argmax(dim=-1) == gt_class).sum() / gt_class.shape[0]


class Model_Inference(nn.Module):
    def __init__(self, net_config, extract_proposals=False, eps=1.4, compute_adj_mat_from_links=False):
        super().__init__()
        self.extract_proposals = extract_proposals
        self.reg_mu = net_config.reg_mu
        self.reg_sigma = net_config.reg_sigma
        c = net_config
        common = dict(activation=c.activation, norm_layer=c.norm_layer, num_groups=c.num_groups)
        conv_out = c.graph_convolution_stem_channels[-1]
        self.encode_node_feat = graph_feature_encoding(in_channels=c.input_node_feat_dim,
                                                       stem_channels=c.node_feat_enc_stem_channels, **common)
        self.encode_edge_feat = graph_feature_encoding(in_channels=c.input_edge_feat_dim,
                                                       stem_channels=c.edge_feat_enc_stem_channels, **common)
        self.pass_messages = graph_convolution(in_node_channels=c.node_feat_enc_stem_channels[-1],
                                               in_edge_channels=c.edge_feat_enc_stem_channels[-1],
                                               stem_channels=c.graph_convolution_stem_channels,
                                               msg_mlp_hidden_dim=c.msg_mlp_hidden_dim,
                                               aggregation=c.aggregation, **common)
        self.predict_node = node_segmentation(in_channels=conv_out, stem_channels=c.node_pred_stem_channels,
                                              num_classes=c.num_classes, **common)
        self.predict_offset = node_offset_predictions(in_channels=conv_out, stem_channels=c.node_pred_stem_channels,
                                                      reg_offset_dim=c.reg_offset_dim, **common)
        self.predict_link = link_predictions(in_channels=conv_out, num_blks_for_edges=c.num_blocks_to_compute_edge,
                                             stem_channels=c.link_pred_stem_channels, num_classes=c.num_edge_classes,
                                             **common)
        self.predict_class = object_classification(in_channels=conv_out, stem_channels=c.node_pred_stem_channels,
                                                   num_classes=c.num_classes, **common)
        if extract_proposals:
            self.set_param_for_proposal_extraction(eps, compute_adj_mat_from_links)

    # ---- reference utility methods (gnn_detector.py:122-139) ----
    @staticmethod
    def freeze_weights(nn_module):
        for p in nn_module.parameters():
            p.requires_grad = False
        return nn_module

    def freeze_layers_except_object_class_predictor(self):
        for name in ('encode_node_feat', 'encode_edge_feat', 'pass_messages', 'predict_node', 'predict_offset',
                     'predict_link'):
            self.freeze_weights(getattr(self, name))

    def set_param_for_proposal_extraction(self, eps, compute_adj_mat_from_links):
        self.compute_adj_mat_from_links = compute_adj_mat_from_links
        self.extract_proposals = True
        self.meas_noise_cov = 0.5 * np.eye(2, dtype=np.float32)
        self.clustering_obj = Simple_DBSCAN(eps, compute_adj_mat_from_links)

    # ---- batched entry used by Model_Training and the benchmarks ----
    def forward_batch(self, gb: GraphBatch, node_features: torch.Tensor, edge_features: torch.Tensor,
                      training: Optional[bool] = None):
        """All frames of `gb` at once.  Returns (node_cls, node_off, link_cls, obj_cls) for the whole batch."""
        return run_detector(self, gb, node_features, edge_features, training)

    def forward(self, node_features: torch.Tensor, edge_features: torch.Tensor, edge_index: torch.Tensor,
                adj_matrix: Optional[torch.Tensor] = None, cluster_node_idx: Optional[List[torch.Tensor]] = None,
                other_features: Optional[torch.Tensor] = None, augmented_features: Optional[torch.Tensor] = None):
        if augmented_features is not None:
            raise NotImplementedError('augmented_features are not used by the reference model')
        n = node_features.shape[0]
        gb = GraphBatch.from_edge_index(edge_index, n)
        cluster_members_list = None
        if cluster_node_idx is not None:
            gb.set_clusters([cluster_node_idx], [0], node_features.device)
            node_cls, node_off, link_cls, obj_cls = run_detector(self, gb, node_features, edge_features)
        else:
            # proposal extraction (gnn_detector.py:164-187): offsets -> predicted centres -> connected components of the
            # predicted links / of the eps graph (on the device, csrc/rgnn_cluster.cu) -> class head on the found clusters
            if not hasattr(self, 'clustering_obj'):
                raise AttributeError("call set_param_for_proposal_extraction(eps, ...) before forward without "
                                     "cluster_node_idx (the reference fails here too: gnn_detector.py:170)")
            gb.set_clusters([[]], [0], node_features.device)
            found = {}

            def cluster(off0, link0):
                reg = unnormalize_gt_offsets(off0.clone(), self.reg_mu, self.reg_sigma)
                centres = other_features[:, :2].to(torch.float32) + reg
                res = self.clustering_obj.cluster_nodes_device(centres, gb, link0)
                found['members'] = res.member_lists()
                gb.cl_ptr, gb.cl_members, gb.n_clusters = res.cl_ptr, res.cl_members, res.n_clusters
                gb.frame_cluster_ptr = [0, res.n_clusters]

            needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
            if not needs_grad:
                # inference: ONE pass; the class head alone runs after the clustering (on the stem output kept in the workspace)
                node_cls, node_off, link_cls, obj_cls = run_detector_then_cluster(self, gb, node_features, edge_features, cluster)
            else:
                # fine-tuning of the class head (Model_Object_Classifier_Finetuning): a first pass without gradient finds the
                # clusters, the second one is the differentiable forward
                with torch.no_grad():
                    _, off0, link0, _ = run_detector(self, gb, node_features, edge_features, training=False)
                cluster(off0, link0)
                node_cls, node_off, link_cls, obj_cls = run_detector(self, gb, node_features, edge_features)
            cluster_members_list = found['members']
        if self.extract_proposals:
            return node_cls, node_off, link_cls, obj_cls, cluster_members_list
        return node_cls, node_off, link_cls, obj_cls


class Model_Inference_v1(nn.Module):
    """The reference's variant with ONE stem for the node-class and offset heads (`node_predictions`; reference
    gnn_detector.py:204-312).  Composed from the block-level modules, each of which runs its fused kernel (inference forward;
    the reference ships no checkpoint or training loop for this variant).  `adj_matrix` may be the dense (N, N) bool matrix
    of the reference or None, in which case the undirected links are taken from `edge_index`."""

    def __init__(self, net_config):
        super().__init__()
        c = net_config
        common = dict(activation=c.activation, norm_layer=c.norm_layer, num_groups=c.num_groups)
        conv_out = c.graph_convolution_stem_channels[-1]
        self.encode_node_feat = graph_feature_encoding(in_channels=c.input_node_feat_dim,
                                                       stem_channels=c.node_feat_enc_stem_channels, **common)
        self.encode_edge_feat = graph_feature_encoding(in_channels=c.input_edge_feat_dim,
                                                       stem_channels=c.edge_feat_enc_stem_channels, **common)
        self.pass_messages = graph_convolution(in_node_channels=c.node_feat_enc_stem_channels[-1],
                                               in_edge_channels=c.edge_feat_enc_stem_channels[-1],
                                               stem_channels=c.graph_convolution_stem_channels,
                                               msg_mlp_hidden_dim=c.msg_mlp_hidden_dim,
                                               aggregation=c.aggregation, **common)
        self.predict_node = node_predictions(in_channels=conv_out, stem_channels=c.node_pred_stem_channels,
                                             num_classes=c.num_classes, reg_offset_dim=c.reg_offset_dim, **common)
        self.predict_link = link_predictions(in_channels=conv_out, num_blks_for_edges=c.num_blocks_to_compute_edge,
                                             stem_channels=c.link_pred_stem_channels, num_classes=c.num_edge_classes, **common)
        self.predict_class = object_classification(in_channels=conv_out, stem_channels=c.node_pred_stem_channels,
                                                   num_classes=c.num_classes, **common)

    def forward(self, node_features: torch.Tensor, edge_features: torch.Tensor, edge_index: torch.Tensor,
                adj_matrix: Optional[torch.Tensor], cluster_node_idx: List[torch.Tensor],
                augmented_features: Optional[torch.Tensor] = None):
        if augmented_features is not None:
            raise NotImplementedError('augmented_features are not used by the reference model')
        x = self.encode_node_feat(node_features)
        e = self.encode_edge_feat(edge_features)
        x = self.pass_messages(x, e, edge_index)
        node_cls, node_off = self.predict_node(x)
        link_cls = self.predict_link(x, adj_matrix if adj_matrix is not None else edge_index)
        obj_cls = self.predict_class(x, cluster_node_idx)
        return node_cls, node_off, link_cls, obj_cls


class Model_Training(nn.Module):
    def __init__(self, net_config, device):
        super().__init__()
        self.pred = Model_Inference(net_config)
        self.loss = Loss_Graph(net_config, device)
        self.device = device
        self.offset_mu = net_config.offset_mu
        self.offset_sigma = net_config.offset_sigma

    @staticmethod
    def pack_batch(node_features: List[torch.Tensor], edge_features: List[torch.Tensor],
                   edge_index: List[torch.Tensor], cluster_node_idx: List[List[torch.Tensor]]):
        """Lists of per-frame tensors (reference collate_fn, datagen_gnn.py:143-190) -> one block-diagonal batch."""
        n_list = [int(x.shape[0]) for x in node_features]
        gb = GraphBatch.from_frames(edge_index, n_list)
        gb.set_clusters(cluster_node_idx, gb.frame_node_ptr[:-1], node_features[0].device)
        nf = node_features[0] if len(node_features) == 1 else torch.cat(node_features, dim=0)
        ef = edge_features[0] if len(edge_features) == 1 else torch.cat(edge_features, dim=0)
        return gb, nf, ef

    def forward(self, node_features: List[torch.Tensor], edge_features: List[torch.Tensor],
                edge_index: List[torch.Tensor], adj_matrix: List[torch.Tensor], labels: Dict[str, List[torch.Tensor]]):
        gb, nf, ef = self.pack_batch(node_features, edge_features, edge_index, labels['cluster_node_idx'])
        return self.forward_packed(gb, nf, ef, labels)

    def forward_packed(self, gb: GraphBatch, nf: torch.Tensor, ef: torch.Tensor, labels: Dict[str, List[torch.Tensor]],
                       global_counts=None, dp_tail: Optional[torch.Tensor] = None):
        """global_counts / dp_tail are supplied by DataParallelTrainer.step only (counts summed over the ranks, NaN flag +
        loss shares that ride in the gradient all-reduce); every other caller divides by this batch's own counts, as
        the reference does (gnn/loss.py:58-70)."""
        node_cls, node_off, link_cls, obj_cls = run_detector(self.pred, gb, nf, ef)
        cat = lambda v: v if isinstance(v, torch.Tensor) else torch.concat(v, dim=0)
        obj_gt, edge_gt, node_gt = cat(labels['cluster_labels']), cat(labels['edge_class']), cat(labels['node_class'])
        # Like the reference (gnn_detector.py:463-465) the in-place normalize_gt_offsets acts on the FRESH tensor that
        # torch.concat returns, so the caller's per-frame label tensors keep their values; an already concatenated tensor
        # (this library's packed fast path) is cloned for the same reason.
        off = labels['node_offsets']
        off = off.clone() if isinstance(off, torch.Tensor) else torch.concat(off, dim=0)
        off_gt = normalize_gt_offsets(off, self.offset_mu, self.offset_sigma)
        loss = self.loss((node_cls, node_off, link_cls, obj_cls), (node_gt, off_gt, edge_gt, obj_gt), global_counts, dp_tail)
        correct = self.loss.last_correct.to(torch.float32)
        accuracy = {'segment_accuracy': correct[0] / node_gt.shape[0],
                    'edge_accuracy': correct[1] / max(edge_gt.shape[0], 1),
                    'object_accuracy': correct[2] / max(obj_gt.shape[0], 1)}
        return loss, accuracy


class Model_Object_Classifier_Finetuning(nn.Module):
    """Fine-tune only the object-class head on DBSCAN proposals (reference gnn_detector.py:481-522)."""

    def __init__(self, net_config):
        super().__init__()
        self.pred = Model_Inference(net_config, extract_proposals=True, eps=net_config.clustering_eps)
        self.loss = Loss_Object_Class(net_config)

    def forward(self, node_features, edge_features, other_features, edge_index, adj_matrix, node_class_labels):
        gts, preds = [], []
        for nf, ef, of, ei, adj, gt in zip(node_features, edge_features, other_features, edge_index, adj_matrix,
                                           node_class_labels):
            _, _, _, obj, members = self.pred(node_features=nf, edge_features=ef, other_features=of,
                                              edge_index=ei, adj_matrix=adj)
            preds.append(obj)
            for m in members:
                gts.append(torch.argmax(torch.bincount(gt[m])))
        obj_gt = torch.stack(gts, dim=0)
        obj_pred = torch.concat(preds, dim=0)
        return self.loss(obj_pred, obj_gt), compute_accuracy(obj_pred, obj_gt)
