"""roar_b200: B200-native (sm_100a) TTS supplementary-data extraction for Roar.

Drop-in for the hot path of ``scripts/dataset_processing/tts/extract_sup_data.py`` /
``TTSDataset`` / ``FilterbankFeatures`` of AshwinSankar17/Roar: log-mel, pYIN pitch,
frame energy, beta-binomial alignment prior, corpus pitch statistics.  All arithmetic
runs in hand-written CUDA behind the C-ABI declared in ``include/roar_sup.h``; there is
no CPU fallback (importing the compute API without the built library raises).
"""
__version__ = "0.1.0"
