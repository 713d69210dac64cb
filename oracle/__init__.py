"""CPU oracle for the NELE-GAN intelligibility-labelling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the CPU arm that is timed *beside* the GPU engine.  The product path
(``nele_gan_b200``) never imports this package and fails loudly when the CUDA
library is missing.

What is pinned and what is not
------------------------------
* HASPI v1/v2 (``haspi_np``): restatement of ``pyHASPI/pyhaspi2.py``.  PINNED:
  ``tests/golden/make_golden.py`` imports the *unmodified* reference module in
  the build container (behind ``oracle/librosa_shim``) and the committed
  fixtures ``tests/golden/*.npz`` hold its outputs; ``tests/test_oracle_haspi.py``
  checks the restatement against them.
* ``librosa.resample`` -> resampy ``kaiser_best`` (``resampy_kaiser``): the
  package is absent from the reference tree and from this image (no network).
  Restated from the published algorithm (resampy 0.2.x ``interpn.resample_f``
  + ``sinc_window(num_zeros=64, precision=9, rolloff=0.9475937167399596,
  kaiser beta=14.769656459379492)``).  PARITY UNPINNED for this sub-step.
* ESTOI (``pystoi_np``): pystoi is an un-vendored, un-pinned pip dependency
  (README.md:14, call sites intel.py:126,133).  Restated from the published
  algorithm (Jensen & Taal 2016; pystoi 0.3.x).  PARITY UNPINNED.
* SIIB / SIIB^Gauss (``pysiib_np``): pysiib is un-vendored and un-pinned
  (README.md:13, call sites intel.py:77,100).  Restated from Van Kuyk et al.
  2018; anchored on the helper copies the reference keeps in-tree
  (intel.py:16-54 framing/get_vad/stft).  PARITY UNPINNED.
"""
