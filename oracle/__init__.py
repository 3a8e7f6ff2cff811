"""CPU oracle for the camera-ISP hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``taichi_image_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Contents
--------
``isp_oracle.py``   numpy float32 restatement of the reference kernels
                    (uc-vision/taichi_image 0.3.2), every function citing the
                    reference file:line it follows.
``c/isp_oracle.c``  plain-C (OpenMP) restatement of the fused packed12 -> RGB
                    path, used as the timed multi-core CPU baseline.
``taichi_shim/``    a minimal pure-Python emulation of the Taichi API surface
                    the reference uses; it lets the UNMODIFIED reference
                    sources under /root/reference execute in this container
                    (Taichi itself is not installable: no network, unpinned).
``gen_golden.py``   runs the reference through the shim and writes the golden
                    vectors under ``tests/golden/``.

Parity pinning status
---------------------
* packed encode/decode: pinned by the reference's only asserting test
  (test/packed.py:6-15, random 12-bit round trip) and by golden vectors
  produced by executing the reference source through ``taichi_shim``.
* demosaic / tone map / metering / resize / transform / ISP: the reference's
  own tests hold NO golden vectors.  They are pinned
  here by golden vectors generated from the reference's unmodified source
  executed on ``taichi_shim`` (Taichi semantics restated: f32 default_fp,
  truncating casts, round-half-away ``ti.round``, RNE f16 stores).  The real
  Taichi runtime could not be executed, so this is "pinned to the reference
  source under emulated Taichi semantics", not to Taichi's code generator.
"""
