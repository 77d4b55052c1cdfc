"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (torch-CPU fp32 / numpy / plain Python) of the reference's
algorithms for the VAD hot path, each function citing the reference file:line it
follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package, and there only as the checker or
the CPU baseline -- never from the product package
(voice-activity-detection-vad-onnx_b200/, importable as `vadx`).

Pinning: oracle/make_golden.py imports the REAL reference nn.Modules and
post-processing classes from /root/reference (AST-extracted, never copied),
loads the same seeded weights into them and freezes their outputs under
tests/golden/.  tests/test_oracle_*.py check these restatements against those
frozen outputs.  Networks that are NOT in /root/reference (Silero, the NeMo
MarbleNet encoder, the modelscope mask-net shell) say "parity unpinned" in
their module headers.
"""
