"""CPU oracle for the EgoHMR diffusion-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker (or as the thing timed for the CPU baseline), never as a fallback for the CUDA library.

It is a plain numpy restatement of the reference's algorithm (sanweiliti/EgoHMR @ dc4e0a0); every function cites the
reference ``file:line`` it follows.  Pinning status:

* schedule / sampler / denoiser / geometry (everything that lives under ``/root/reference``): PINNED against the
  reference's own PyTorch code imported in the build container — ``tests/golden/make_golden.py`` runs the unmodified
  reference on seeded synthetic inputs and commits the vectors under ``tests/golden/``; ``tests/test_oracle_golden.py``
  checks this oracle against them.
* SMPL forward (``oracle/smpl.py``): the arithmetic lives in the un-vendored third-party dependency ``smplx==0.1.28``
  (environment.yml:197) which is not installed anywhere we can reach, and the reference holds no golden vectors for
  it => **parity unpinned** for that component: it restates the published algorithm of ``smplx/lbs.py::lbs`` and is
  anchored only on the reference's call sites and on known-answer tests we author (identity pose, rigid root rotation).
* collision guidance: COAP (unpinned git dependency) is not reproducible offline; the guided sampler is pinned only
  at the boundary, with a synthetic differentiable collision callable plugged into both sides.
"""
