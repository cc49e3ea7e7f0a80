"""kmercamel_b200 — B200-native implementation of KmerCamel's `compute` hot path.

The product is the C-ABI library csrc/libkcgpu.so (CUDA, sm_100a) plus the `kmercamel` CLI in host/.
This package is only the ctypes binding used by tests and bench.py; it never falls back to a CPU path.
"""
from .api import (Context, ComputeResult, Group, KcError, frame_fasta, frame_fasta_file, group_plan, limbs_for_k, lib_path,
                  load_library)

__all__ = ["Context", "ComputeResult", "Group", "KcError", "frame_fasta", "frame_fasta_file", "group_plan", "limbs_for_k", "lib_path",
           "load_library"]
