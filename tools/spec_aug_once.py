#!/usr/bin/env python3
"""Runs SpecAugment on the bench shape a few times (for ncu captures): python tools/spec_aug_once.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200
import bench
print(bench.spec_aug_microbench(asr_b200, torch.device("cuda"), iters=3))
