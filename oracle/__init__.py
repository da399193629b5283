"""CPU oracle for the image-augmentation hot path.  TEST INFRASTRUCTURE ONLY.

  skimage_oracle.py  part A: numpy/scipy restatement of the scikit-image calls
                     the reference's tests use (tests/millipyde_tests.py).
  ref_exact.c/.py    part B: plain-C restatement of the reference's own kernels
                     on its RGBA8 / fp64 layouts (src/millipyde_image.cpp).
  build_ref.py       part C: recipe that compiles the UNMODIFIED reference from
                     /root/reference through a HIP->CUDA macro shim into
                     oracle/_ref/ (needs a GPU to import; used under gpurun).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product
(millipyde_b200/) never does and has no CPU fallback.
"""
