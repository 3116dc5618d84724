#!/bin/bash
# round 2, first GPU pass of the tile engine: A/B bit-identity, the full GPU suite, then an engine A/B on the bench
python -m pytest tests/test_gpu_parity.py -x -q -k "tile_pair_engine" 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
tools/ab_env.sh SMD_PAIR_ENGINE=0 SMD_PAIR_ENGINE=1 "SMD_PAIR_ENGINE=1 SMD_XSUB=1" "SMD_PAIR_ENGINE=1 SMD_XSUB=2"
