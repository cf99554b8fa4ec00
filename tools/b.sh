#!/bin/bash
# build librgnn.so from anywhere
cd "$(dirname "$0")/.." && python -m graph_neural_network_for_radar_perception_b200.build 2>&1 | grep -E "error|Error|warning: variable" | head -20
