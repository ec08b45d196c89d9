"""CPU oracle for the localization-and-crop hot path of spine-vision.

TEST INFRASTRUCTURE ONLY.  Nothing under ``spine_vision_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline -- never as the product path.

Two layers live here:

* ``oracle.reference_path`` -- a *port* of the reference's own Python for the
  path (``spine_vision/io/__init__.py:15-30``,
  ``spine_vision/datasets/classification/cropping.py:104-169, 316-483``,
  ``spine_vision/training/datasets/classification.py:40-68, 247-278``) that
  calls the same third-party libraries the reference calls (NumPy, Pillow,
  OpenCV, torch).  This is what ``bench.py --impl reference`` times.
* ``oracle.fixedpoint`` / ``oracle.convnext`` -- restatements of the
  third-party arithmetic that is not under ``/root/reference``: Pillow's 8-bit
  antialiased BILINEAR resize (pillow 10.2.0 in the reference's ``uv.lock``;
  12.2.0 installed), OpenCV's 8U ``INTER_LINEAR`` resize (opencv-python
  4.6.0.66 locked; 4.13.0 installed) and timm 1.0.22's ConvNeXt forward (timm
  is not installed).  These state, in integer / fp32 arithmetic, exactly what
  the CUDA kernels implement.

* ``oracle.itk_resample`` / ``oracle.metaimage`` / ``oracle.dicom`` -- restatements of what SimpleITK 2.5.3 (ITK, MetaIO,
  GDCM; not installed) does in front of the path: the 0.3 mm resample + LPI orientation + middle slice, MetaImage reading,
  DICOM series reading.  **Parity unpinned** (each header says so): they restate documented conventions and are what
  K0 and the native decoders are tested against.
* ``oracle.make_golden`` / ``make_golden_host`` / ``make_golden_localization`` -- container-only scripts that run the
  reference's OWN functions and OWN dataset builders (through ``oracle.ref_shim``) and freeze their outputs into
  ``tests/golden/``.

Pinning status: the reference ships **no tests and no golden vectors** for
this path (``AGENTS.md:552-554``).  The oracle is pinned instead against
outputs of the reference's own unmodified functions, imported in the build
container through ``oracle/ref_shim.py`` and frozen into ``tests/golden/`` by
``oracle/make_golden.py`` (committed), plus the one exact known answer the
reference's notebooks record (``mm_to_pixels``,
``notebooks/compare_crop_modes.ipynb:65,255``).  The ConvNeXt arithmetic itself
(timm) and the SimpleITK resample are absent from the container, so for those
two pieces parity is "unpinned" beyond structural equivalence with
``torchvision.models.convnext_base``.
"""
