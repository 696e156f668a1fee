"""Parity oracle for the FB-BEV view-transformation hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; the product (``fb-bev_b200/``) never does, and fails loudly when its
CUDA library is missing instead of falling back to anything in here.

* ``oracle.cpu``      -- ctypes binding of ``oracle/libfbbev_oracle.so`` (the C
  restatement in ``fbbev_oracle.c``) plus numpy glue.
* ``oracle.torch_ref`` -- float32 PyTorch-CPU restatement of the Python side of
  the path (geometry, depth-aware spatial cross-attention, encoder layer).
* ``oracle.ref_cuda``  -- ctypes binding of ``oracle/_ref/libbev_pool_ref.so``:
  the reference's own ``bev_pool_cuda.cu`` compiled for sm_100a (GPU box only).
"""
