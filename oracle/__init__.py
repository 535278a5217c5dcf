"""CPU oracle for the GAP/ADMM-TV hot path of PnP_SCI/python.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status
-------------
* R1-R5, R7, R8, R10 (``A_``, ``At_``, ``psnr``, ``gap_denoise``,
  ``admm_denoise``, ``admmdenoise_cacti``, ``gap_denoise_bayer``): restated in
  ``oracle/pnp_sci.py`` and PINNED against the reference's own code, imported
  unmodified from /root/reference in the build container by
  ``oracle/reference_loader.py`` (fixtures: ``tests/golden/*.npz``, generator:
  ``tests/golden/make_golden.py``).
* R6 (``skimage.restoration.denoise_tv_chambolle``): scikit-image is a
  third-party dependency that is neither vendored in the reference tree nor
  installed here (``PnP_SCI/python/environment.yml:16``, unpinned;
  ``pnp_sci_algo.py:13-17`` only works with skimage < 0.18).  ``oracle/
  tv_chambolle.py`` restates the published algorithm of scikit-image 0.17.2
  ``skimage/restoration/_denoise.py::_denoise_tv_chambolle_nd``.  The
  reference holds no runnable test or golden vector at this boundary, so this
  one function is **parity unpinned**; it is corroborated (interior pixels, to
  1e-15 in float64) by a transliteration of the in-tree MATLAB
  ``tvdenoise_cham_ITV2D.m`` (``tests/test_oracle_tv.py``).
* Joint module: ``joint_admm_denoise``, ``gap_multistep_denoise``,
  ``admm_multistep_denoise``, ``gap_joint_denoise``, ``admm_joint_denoise`` in
  ``oracle/pnp_sci.py``: PINNED against ``joint_pnp_sci_algo.py`` run unmodified, a fixed
  elementwise stand-in injected in FFDNet's place for the TV+CNN periods.
* IQA (``oracle/iqa.py``: ``compare_psnr``, ``compare_ssim`` of scikit-image 0.17.2) and the
  MATLAB twin's ``TV_denoising.m`` (``oracle/matlab_tv.py``): restated from the published /
  in-tree sources, **parity unpinned** (scikit-image, MATLAB and Octave are absent here);
  checked by properties in ``tests/``.
"""
