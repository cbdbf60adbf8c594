"""B200-native meteor-scatter detection hot path (drop-in for th-nuernberg/meteor-scatter's detectors).

Layers (host code is Python; all sample arithmetic runs in csrc/ CUDA kernels behind include/ms_b200.h):

* drop-in modules mirroring the reference's layout:
  ``dsp.src.main.proc_wav_file``, ``dsp.src.live.backend.processor.wav_file_process`` (+ ``aggregates``),
  ``meteor_detect_class.prime_detection.plot_spectrogram`` / ``detector_and_classification``;
* batch API: ``pipeline.DetectorA`` (``run``, ``run_pass``, ``run_host``, ``capture``), ``pipeline.PassPipeline``,
  ``batch.process_files``; ``csvout`` writes the dashboard's ``YYYYMMDD.csv`` day files; ``wavio`` reads WAVs;
* ``ops``: thin torch-facing wrappers over the C-ABI; ``_lib``: the ctypes binding; ``build``: nvcc recipe.

Importing this package does not load the shared library; the first compute call does and raises if it has not been
built (``python -c "import __graft_entry__ as g; g.build()"``).  There is no CPU fallback.
"""

__version__ = "0.1.0"
__all__ = ["batch", "build", "csvout", "ops", "pipeline", "synth", "wavio"]
