"""ctypes bindings of shredword_b200/libtrainer.so.

Mirrors reference shredword/cbase.py: the same library discovery (a file starting with `trainer` /
`libtrainer` next to this module, cbase.py:5-19), the same structure layouts (cbase.py:36-42) and the
same eight signatures (cbase.py:44-59), followed by the additive entry points of
include/shredword_b200.h. There is no Python or CPU fallback: if the CUDA library is missing, import
fails; if there is no GPU, the compute calls fail with the library's error message.
"""
import ctypes, os
from ctypes import Structure, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_size_t, c_char_p, c_void_p, c_double, POINTER

def _get_lib_path():
  pkg_dir = os.path.dirname(os.path.abspath(__file__))
  for search_dir in (pkg_dir, os.path.join(pkg_dir, 'lib')):
    if not os.path.isdir(search_dir): continue
    for file in sorted(os.listdir(search_dir)):
      if (file.startswith('trainer') or file.startswith('libtrainer')) and file.endswith('.so'):
        return os.path.join(search_dir, file)
  raise FileNotFoundError(
    f"Could not find libtrainer.so in {pkg_dir}: build it with `python -m shredword_b200.build` "
    "(needs nvcc; there is no CPU fallback)")

lib = ctypes.CDLL(_get_lib_path())

MIN_HEAP_SIZE = 4096
INITIAL_VOCAB_SIZE = 256

class PairKey(Structure): _fields_ = [("first", c_int32), ("second", c_int32)]
class HeapEntry(Structure): _fields_ = [("key", PairKey), ("freq", c_uint64), ("version", c_uint32)]
class MaxHeap(Structure): _fields_ = [("data", POINTER(HeapEntry)), ("size", c_size_t), ("cap", c_size_t)]
class Corpus(Structure): _fields_ = [("words", c_void_p), ("word_counts", POINTER(c_uint64)), ("vocab_size", c_size_t)]
class BIMap(Structure): _fields_ = [("buckets", c_void_p), ("nbuckets", c_size_t)]
class BPEConfig(Structure):
  _fields_ = [("target_vocab_size", c_size_t), ("unk_id", c_int32), ("character_coverage", c_float), ("min_pair_freq", c_uint64)]
class Trainer(Structure):
  _fields_ = [("config", BPEConfig), ("heap", MaxHeap), ("corpus", Corpus), ("bigram_map", BIMap), ("next_token", c_size_t),
              ("num_merges", c_size_t), ("merge_ops", POINTER(PairKey)), ("token_strs", POINTER(c_char_p)),
              ("token_freq", POINTER(c_uint64)), ("impl", c_void_p)]
class SwbStats(Structure):
  _fields_ = [("kernel_launches", c_uint64), ("merge_launches", c_uint64), ("load_ms", c_double), ("count_ms", c_double),
              ("merge_ms", c_double), ("merge_kernel_ms", c_double), ("merge_scan_bytes", c_uint64), ("merge_alg_bytes", c_uint64),
              ("rows", c_uint64), ("live_symbols", c_uint64), ("words", c_uint64), ("long_words", c_uint64), ("repacks", c_uint64),
              ("host_pop_ms", c_double), ("host_launch_ms", c_double), ("host_wait_ms", c_double), ("host_apply_ms", c_double),
              ("records", c_uint64), ("heap_pushes", c_uint64), ("heap_pops", c_uint64), ("heap_peak", c_uint64),
              ("collectives", c_uint64), ("exchange_bytes", c_uint64),
              ("resident_spill_merges", c_uint64), ("exchange_ns", c_uint64), ("reserved_", c_uint64 * 3),
              ("resident_local_merges", c_uint64), ("resident_grid_merges", c_uint64),
              ("resident_local_ms", c_double), ("resident_grid_ms", c_double),
              ("hints_sent", c_uint64), ("hints_taken", c_uint64), ("hints_rejected", c_uint64), ("host_peek_ms", c_double),
              ("tokenize_ms", c_double), ("tokenize_bytes", c_uint64),
              ("local_by_log", c_uint64 * 4), ("local_by_log_ms", c_double * 4), ("local_by_log_recs", c_uint64 * 4)]

# ---- the reference's eight entry points (reference cbase.py:44-59)
lib.create_trainer.argtypes = [POINTER(BPEConfig)]
lib.create_trainer.restype = POINTER(Trainer)
lib.bpe_trainer_destroy.argtypes = [POINTER(Trainer)]
lib.bpe_trainer_destroy.restype = None
lib.bpe_init.argtypes = [POINTER(Trainer)]
lib.bpe_init.restype = None
lib.bpe_count_bigrams.argtypes = [POINTER(Trainer)]
lib.bpe_count_bigrams.restype = None
lib.bpe_load_corpus.argtypes = [POINTER(Trainer), c_char_p]
lib.bpe_load_corpus.restype = c_int
lib.bpe_merge_batch.argtypes = [POINTER(Trainer), c_int]
lib.bpe_merge_batch.restype = c_int
lib.bpe_train.argtypes = [POINTER(Trainer)]
lib.bpe_train.restype = c_int
lib.bpe_save.argtypes = [POINTER(Trainer), c_char_p, c_char_p]
lib.bpe_save.restype = None

# ---- additive entry points (include/shredword_b200.h, part 2)
T = POINTER(Trainer)
_sigs = {
  "swb_last_error": ([], c_char_p),
  "swb_set_log_level": ([c_int], None),
  "swb_device_count": ([], c_int),
  "swb_set_device": ([c_int], c_int),
  "swb_release_cached_memory": ([], c_size_t),
  "swb_load_corpus_buffer": ([T, c_void_p, c_size_t], c_int),
  "swb_load_corpus_device": ([T, c_void_p, c_size_t], c_int),
  "swb_num_merges": ([T], c_size_t),
  "swb_get_merges": ([T, c_void_p, c_size_t], c_size_t),
  "swb_token_bytes": ([T, c_int32, c_void_p, c_size_t], c_size_t),
  "swb_get_byte_map": ([T, c_void_p], None),
  "swb_token_freq": ([T, c_void_p, c_size_t], c_int),
  "swb_num_words": ([T], c_size_t),
  "swb_num_symbols": ([T], c_size_t),
  "swb_word_bytes_total": ([T], c_size_t),
  "swb_get_words": ([T, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
  "swb_get_stats": ([T, POINTER(SwbStats)], None),
  "swb_set_kernel_timing": ([T, c_int], None),
  "swb_profile_scripted_merges": ([T, c_void_p, c_size_t, POINTER(c_double)], c_int),
  "swb_encoder_create": ([c_void_p, c_size_t, c_void_p], c_void_p),
  "swb_encoder_from_trainer": ([T], c_void_p),
  "swb_encoder_destroy": ([c_void_p], None),
  "swb_encode": ([c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)], c_int64),
  "swb_encode_device": ([c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)], c_int64),
  "swb_decode": ([c_void_p, c_void_p, c_size_t, c_void_p, c_size_t], c_size_t),
  "swb_encoder_kernel_launches": ([c_void_p], c_uint64),
  "swb_normalize": ([c_void_p, c_size_t, c_void_p, c_size_t, c_int], c_int64),
  "swb_pretokenize": ([c_void_p, c_size_t, c_void_p, c_size_t, c_int], c_int64),
  "swb_set_shard": ([T, c_int, c_int], c_int),
  "swb_device_pci_bus_id": ([c_int, c_char_p, c_size_t], c_int),
  "swb_dist_reduce_records": ([c_void_p, c_size_t], c_size_t),
  "swb_dist_seed": ([T, c_void_p, c_size_t], None),
  "swb_dist_next_merge": ([T, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)], c_int),
  "swb_dist_apply": ([T, c_void_p, c_size_t], None),
  "swb_dist_peek_next": ([T, POINTER(c_int32), POINTER(c_int32), POINTER(c_uint64)], c_int),
  "swb_dist_peek_list": ([T, c_void_p, c_size_t], c_size_t),
  "swb_shard_count": ([T, c_void_p, c_size_t], c_int64),
  "swb_shard_merge": ([T, c_int32, c_int32, c_int32, c_void_p, c_size_t], c_int64),
  "swb_dist_unique_id": ([c_void_p], c_int),
  "swb_dist_init": ([T, c_int, c_int, c_void_p], c_int),
  "swb_dist_set_sharded": ([T, c_int], c_int),
  "swb_load_corpus_shard": ([T, c_void_p, c_size_t, c_uint64, c_int], c_int),
  "swb_dist_has_comm": ([c_int, c_int], c_int),
  "swb_dist_shutdown": ([], None),
  "swb_save_with_freq": ([T, c_char_p, c_char_p, c_void_p, c_size_t], c_int),
}
for _name, (_args, _res) in _sigs.items():
  getattr(lib, _name).argtypes = _args
  getattr(lib, _name).restype = _res

def last_error() -> str:
  return (lib.swb_last_error() or b"").decode("utf-8", "replace")
