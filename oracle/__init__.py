"""CPU oracle for the PonderV2 pretraining hot path — TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
`cpu_baseline` / `--impl reference` leg may import it, and only as the checker / timed CPU baseline.
The product (ponderv2_b200/) never imports this package and has no CPU fallback.

Pinning status:
  * trilinear sampler .... pinned: the reference's own KAT (libs/smooth-sampler/smooth_sampler/modules.py:104-156:
                           allclose vs F.grid_sample fwd + first-order grads, fp64 gradcheck/gradgradcheck) is run
                           against oracle/trilinear_oracle.py in tests/test_oracle_cpu.py.
  * NeuS renderer ........ pinned: oracle/render_oracle.py is checked against golden vectors produced by importing
                           the reference's own ponder/models/ponder/render_utils (oracle/gen_golden.py, fixtures in
                           tests/golden/render_*.npz).
  * sparse convolution ... PARITY UNPINNED: the arithmetic lives in third-party `spconv` (unpinned `spconv-cu113`,
                           README.md:61-63), absent from /root/reference and not installable offline; the reference
                           holds no test or golden vector for it.  oracle/spconv_oracle.py restates the published
                           semantics anchored on the reference's call sites
                           (spconv_unet_v1m1_base.py:47-66,111-119,135-142,171-177).  Two independently written
                           restatements exist — numpy/torch (spconv_oracle.py) and plain C (spconv_c.c, bound by
                           c_oracle.py) — and tests/test_oracle_cpu.py requires them to agree bit-exactly on rulebooks
                           and to 1e-12 on the fp64 convolution; a third anchor is torch's dense conv3d.
"""
