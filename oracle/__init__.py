"""CPU oracle for the ViTok-v2 encode/decode hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy for the
byte/index work, plain torch-CPU tensor ops for the floating-point work) of the
reference algorithm in /root/reference/vitok.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it, and only as the checker or the CPU baseline --
never on the product path.  The product (``vitok-release_b200/``) never imports
this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from /root/reference in the build container, runs it on seeded inputs
and writes the fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against every one of them (bit-exact for
patchify/unpatchify/index packing and decode_variant; <=2e-5 abs on fp32
latents/reconstructions).  The reference ships no golden vectors of its own for
the floating-point path, so the fixtures generated from
the running reference are the pin.
"""
