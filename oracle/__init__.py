"""CPU oracle for the DiffSinger acoustic / variance sampling hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a from-scratch CPU restatement (plain ``torch`` CPU tensor ops, fp32 or
fp64) of the algorithms on the hot path of the reference
(``vsingerxiaoice-rwkv/xiaoicesing-io``, an OpenVPI DiffSinger fork):

* ``oracle.denoisers``  - WaveNet / LYNXNet forward
  (reference ``modules/backbones/wavenet.py:18-107``, ``lynxnet.py:29-163``,
  ``modules/commons/common_layers.py:107-117,266-278``)
* ``oracle.samplers``   - DDPM ancestral / DDIM / PNDM(PLMS) / DPM-Solver++(2M) /
  UniPC(bh2) / rectified-flow Euler, RK2, RK4, RK5
  (reference ``modules/core/ddpm.py:16-351``, ``modules/core/reflow.py:66-138``,
  ``inference/dpm_solver_pytorch.py``, ``inference/uni_pc.py``)
* ``oracle.weights``    - the seeded random-init recipe (SURVEY.md section 8a gotcha 1)

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or the timed CPU baseline.  The product package (``xiaoicesing_io_b200``) never
imports it and has no CPU fallback.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this
path, so the oracle is pinned the other way the task allows:
``oracle/make_golden.py`` imports the *unmodified reference* from
``/root/reference`` (present in the build container only), runs it on seeded
inputs and commits inputs + outputs as small fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against every fixture.
"""
