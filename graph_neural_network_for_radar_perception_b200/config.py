"""Configuration object with the attribute names the reference's model constructors read
(reference modules/set_configurations/set_config_gnn.py:9-114), fed from a YAML file with the reference's
sections and keys (configuration_radarscenes_gnn.yml) or from the built-in defaults below.
"""
from __future__ import annotations

import os

import numpy as np
import yaml

DEFAULT_YML = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'configuration_radarscenes_gnn.yml')

_DYN_CLASSES = ['CAR', 'PEDESTRIAN', 'PEDESTRIAN_GROUP', 'TWO_WHEELER', 'LARGE_VEHICLE', 'NONE', 'FALSE']
_ALL_CLASSES = _DYN_CLASSES + ['STATIC']


def _name_to_id(names):
    return {n: i for i, n in enumerate(names)}


class config:
    def __init__(self, config_filepath: str = DEFAULT_YML):
        with open(config_filepath, 'r') as f:
            y = yaml.safe_load(f)
        sel, grid, arch = y['DATA_SELECTION_PARAM'], y['GRID_LIMITS'], y['GNN_ARCHITECTURE']
        self.seed = y['RANDOM']['seed']
        d = y['DIRECTORIES']
        self.dataset_path, self.model_weights_dir = d['dataset_dir'], d['model_weights_dir']
        self.weights_name, self.finetuned_weights_name = d['weights_name'], d['finetuned_weights_name']
        # accumulation window and graph construction
        self.window_size = sel['temporal_window_size']
        self.ball_query_eps_square = sel['ball_query_eps_square']
        self.k_number_nearest_points = sel['k_number_nearest_points']
        self.reject_static_meas_by_ransac = sel['reject_static_meas_by_ransac']
        self.dataset_augmentation = sel['dataset_augmentation']
        # region of interest
        for k in ('min_x', 'max_x', 'min_y', 'max_y', 'min_sigma_x', 'max_sigma_x', 'min_sigma_y', 'max_sigma_y', 'dx', 'dy'):
            setattr(self, k, grid[k])
        self.grid_min_th, self.grid_min_r = 0, 0
        self.grid_max_th = np.pi * 0.5
        self.grid_max_r = np.sqrt(self.max_x ** 2 + self.max_y ** 2)
        # network
        self.node_features, self.edge_features, self.reg_offset = arch['node_features'], arch['edge_features'], arch['reg_offset']
        self.activation, self.norm_layer, self.num_groups = arch['activation'], arch['normalization'], arch['num_groups']
        self.reg_mu, self.reg_sigma = arch['reg_mu'], arch['reg_sigma']
        self.offset_mu, self.offset_sigma = arch['reg_mu'], arch['reg_sigma']
        self.aggregation = arch['aggregation']
        for k in ('node_feat_enc_stem_channels', 'edge_feat_enc_stem_channels', 'graph_convolution_stem_channels',
                  'msg_mlp_hidden_dim', 'num_blocks_to_compute_edge', 'hidden_node_channels_GAT', 'num_heads_GAT',
                  'link_pred_stem_channels', 'node_pred_stem_channels', 'num_edge_classes'):
            setattr(self, k, arch[k])
        self.input_node_feat_dim, self.input_edge_feat_dim = len(self.node_features), len(self.edge_features)
        self.reg_offset_dim = len(self.reg_offset)
        cats = y['OBJECT_CATEGORIES']
        self.object_classes, self.class_weights = cats['OBJECT_CLASS'], cats['OBJECT_CLASS_WEIGHTS']
        self.object_classes_dyn, self.class_weights_dyn = cats['OBJECT_CLASS_DYN'], cats['OBJECT_CLASS_WEIGHTS_DYN']
        self.num_classes = len(self.object_classes_dyn)
        lw = y['LOSS_WEIGHTS']
        self.edge_cls_loss_weight, self.node_cls_loss_weight = lw['edge_loss_cls'], lw['node_loss_cls']
        self.node_reg_loss_weight, self.obj_cls_loss_weight = lw['node_loss_reg'], lw['obj_loss_cls']
        self.new_labels_to_id_dict = _name_to_id(_ALL_CLASSES)
        self.new_labels_to_id_dict_dyn = _name_to_id(self.object_classes_dyn)
        opt, ft, ds = y['OPTIMIZATION'], y['FINETUNING'], y['DATASET']
        self.optim, self.max_train_iter = opt['optim'], opt['max_training_iterations']
        self.learning_rate, self.weight_decay = opt['learning_rate'], opt['weight_decay']
        self.num_training_samples, self.num_validation_samples = ds['num_training_samples'], ds['num_validation_samples']
        self.shuffle_training_samples, self.shuffle_validation_samples = ds['shuffle_training_samples'], ds['shuffle_validation_samples']
        self.include_region_confidence = y['DATASET_INFO']['include_region_confidence']
        self.optim_finetuning, self.max_train_iter_finetuning = ft['optim'], ft['max_training_iterations']
        self.learning_rate_finetuning, self.weight_decay_finetuning = ft['learning_rate'], ft['weight_decay']
        self.clustering_eps = ft['clustering_eps']
