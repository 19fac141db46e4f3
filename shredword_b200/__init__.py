"""shredword_b200 -- B200-native BPE trainer/encoder hot path behind ShredWord's BPETrainer API.

`from shredword_b200 import BPETrainer` is the drop-in for `from shredword.trainer import BPETrainer`
(reference shredword/__init__.py:1). The CUDA library is required: importing the trainer without
shredword_b200/libtrainer.so raises; there is no CPU fallback. `shredword_b200.synth` (corpus
generator) and `shredword_b200.build` import without it."""

__version__ = '0.1.0'

def __getattr__(name):
  if name in ("BPETrainer", "BPEEncoder"):
    from . import trainer
    return getattr(trainer, name)
  raise AttributeError(name)
