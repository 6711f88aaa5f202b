"""Runs the stored-numerator CE forward + backward a few times at one shape (for ncu captures of single kernels).
usage: python tools/ce_fwd_only.py [N H V]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import ce_probe

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:4]] if len(sys.argv) >= 4 else [12666, 512, 10000]
    ce_probe.run(*a)
