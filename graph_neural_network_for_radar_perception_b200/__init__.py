"""B200-native (sm_100a) implementation of the radar GNN hot path of
UditBhaskar19/GRAPH_NEURAL_NETWORK_FOR_RADAR_PERCEPTION behind the reference's PyTorch-facing API.

    from graph_neural_network_for_radar_perception_b200 import config, Model_Inference, Model_Training
    from graph_neural_network_for_radar_perception_b200.graph_features import compute_adjacency_information

All arithmetic runs in csrc/librgnn.so (C-ABI in include/rgnn.h); there is no CPU fallback.
"""
from .config import config                                                   # noqa: F401
from .gnn_detector import (Model_Inference, Model_Inference_v1, Model_Training,     # noqa: F401
                           Model_Object_Classifier_Finetuning, compute_accuracy)

__all__ = ['config', 'Model_Inference', 'Model_Inference_v1', 'Model_Training', 'Model_Object_Classifier_Finetuning', 'compute_accuracy']
