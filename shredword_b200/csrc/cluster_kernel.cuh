// cluster_kernel.cuh -- the resident merge kernel (default single-GPU path of bpe_merge_batch / bpe_train).
//
// Replaces the scan + splice + delta bookkeeping of reference bpe_merge_batch (reference
// csrc/bpe/bpe.cpp:437-520) for a whole batch of merges with ONE launch that stays resident:
//
//   grid    = thread-block clusters of CL_SIZE CTAs (one CTA per SM), CL_THREADS threads each
//   cluster 0 ("leader") talks to the host: CTA 0 / thread 0 polls a mailbox in mapped host memory for
//             the next pair, the results go back the same way -> one PCIe round trip per merge, no launch
//   LOCAL   a merge whose pair has a birth log of at most CL_LOCAL_MAX entries is done by the leader
//           cluster alone: one thread per candidate word, the signed deltas are aggregated in ONE hash
//           table distributed over the shared memories of the cluster (DSMEM atomics; owner CTA = pair
//           hash), cluster barriers instead of grid-wide atomics + fences, then every CTA applies its
//           own pairs to the device frequency table and writes its records to the host
//   GRID    every other merge (two initial symbols: row-signature scan; very long logs) runs on all
//           clusters through the global pair table, exactly like the per-launch kernel merge_rows
//
// Every spin has a time-out so that a vanished host cannot hang the GPU.
#pragma once

#include <cooperative_groups.h>

#include "merge_kernels.cuh"

namespace swb {

namespace cg = cooperative_groups;

constexpr int CL_THREADS = 512;
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_SIZE = 8;                 // CTAs per cluster (portable maximum)
constexpr int CL_HT_SLOTS = 2048;          // slots per CTA of each delta table (T1: this CTA's partial sums, T2: the pairs this CTA owns)
constexpr int CL_INBOX = 1024;             // partial sums a CTA can receive per merge
constexpr int CL_BIRTH_STAGE = 1024;       // log entries staged per CTA and merge
constexpr int CL_REC_STAGE = 512;          // records staged per CTA and merge
constexpr unsigned int CL_LOCAL_MAX = CL_SIZE * CL_THREADS * 64;  // longest birth log the leader cluster takes alone (entries)
constexpr int CL_CAND_CAP = 1024;          // candidate words listed per CTA and merge (more: rewritten where they are found)
constexpr int CL_MAX_PROBES = 256;
constexpr unsigned int CL_SOLO_MAX = 2 * CL_THREADS;  // SOLO: a log of up to 2 entries (hence at most 2 candidate words) per thread of ONE CTA; measured: at 8 per thread the words of one CTA take longer than the exchange saves
// Effective limits of one launch. The defaults are the compile-time capacities above; the parity tests shrink them
// (SWB_TEST_* environment variables, see TrainerImpl::cluster_tune) so that the overflow paths -- candidate lists past
// the shared-memory list, inbox and table spills into the global pair table, records and log entries past their stages,
// GRID merges from long logs -- run on corpora small enough for the CPU oracle.
struct ClusterTune { unsigned int local_max, cand_cap, inbox_cap, rec_stage, birth_stage, max_probes, solo_max /* longest log CTA 0 of the leader cluster takes all by itself */; };
__host__ __device__ __forceinline__ ClusterTune cluster_tune_default() {
  return ClusterTune{CL_LOCAL_MAX, (unsigned int)CL_CAND_CAP, (unsigned int)CL_INBOX, (unsigned int)CL_REC_STAGE, (unsigned int)CL_BIRTH_STAGE, (unsigned int)CL_MAX_PROBES, (unsigned int)CL_SOLO_MAX};
}
constexpr unsigned int CL_IP_LOCAL_MAX = 8192;  // longest occurrence list of a pair of two initial symbols the leader cluster takes alone (every entry is a candidate word: two per thread)

// host -> device command: ONE 16-byte word in mapped host memory, written with a single 16-byte store and read with a
// single 16-byte load (one PCIe read per poll, no second trip for a payload):
//   x, y = the pair (second, first);  z = new_id | op << 28 (op 0 = merge, 1 = stop);  w = 32-bit check over (seq, x, y, z)
// The sequence number is not carried in the clear: the device knows which one it waits for, and a word that was written
// for any other sequence number (a stale command, an old hint) fails the check. The check is never 0, so a zeroed mailbox
// word matches nothing. Atomicity assumption: a 16-byte aligned store of the host reaches memory as one unit (true for SSE
// stores on every x86-64 with AVX; other hosts write two 8-byte halves, the one with the check word last: HostCmd2Sender::store16 in trainer_impl.cuh);
// a torn word would still have to pass the 32-bit check to be taken.
struct __align__(16) HostCmd2 { unsigned int x, y, z, w; };
__host__ __device__ __forceinline__ unsigned int cmd3_word(unsigned long long seq, unsigned int x, unsigned int y, unsigned int z) {
  unsigned long long f = (seq + 0x9E3779B97F4A7C15ull) * 0xD6E8FEB86659FD93ull;
  f ^= (unsigned long long)x * 0x9E3779B1ull; f = (f << 23 | f >> 41) * 0xA24BAED4963EE407ull;
  f ^= (unsigned long long)y * 0x85EBCA6Bull; f = (f << 29 | f >> 35) * 0x9FB21C651E98DF25ull;
  f ^= (unsigned long long)z * 0xC2B2AE35ull; f ^= f >> 32; f *= 0xD6E8FEB86659FD93ull; f ^= f >> 29;
  const unsigned int w = (unsigned int)f ^ (unsigned int)(f >> 32);
  return w ? w : 0x5BD1E995u;
}
// leader -> other clusters (device memory)
struct DevCmd2 { unsigned long long epoch, pair, new_id_op, log_range, alive_ns, k /* index of the merge within this launch */, t_cmd /* %globaltimer when the command arrived */, spec /* started from a hint */, tail_done /* GRID merges whose tail (emit + publish, run by whichever block finished last) is complete */, done_count /* blocks that have finished their part of a GRID merge, summed over all GRID merges of this launch: never reset, so no reset can race with the next merge's arrivals */, tail_max /* largest frequency the last GRID merge makes the host push (its tail's atomicMax; zeroed with the command) */, tail_flags /* flag word of the header that tail published */; };
__host__ __device__ __forceinline__ unsigned long long cmd2_check(unsigned long long seq, unsigned long long pair,
                                                                  unsigned long long nio, unsigned long long lr) {
  return (seq * HDR_MAGIC) ^ pair ^ (nio << 7 | nio >> 57) ^ (lr * 0xD6E8FEB86659FD93ull);
}
__host__ __device__ __forceinline__ unsigned long long cmd2_check(unsigned long long seq, unsigned long long pair, unsigned long long nio,
                                                                  unsigned long long lr, unsigned long long cursor) {
  return cmd2_check(seq, pair, nio, lr) ^ (cursor * 0xA24BAED4963EE407ull);
}

// per-CTA control block in shared memory
struct ClusterCtl {
  // the command, written into every CTA of the leader cluster by its CTA 0
  unsigned long long pair, new_id_op, log_range, k;
  unsigned int mode, stop;             // mode 0 = LOCAL (the leader cluster), 1 = GRID (all clusters), 2 = SOLO (CTA 0 of the leader cluster alone)
  unsigned int go;                     // CTAs 1.. of the leader cluster: bumped (remote store by CTA 0) when a command is for them too
  unsigned int log_cursor, births_total;  // CTA 0: log entries before this merge / appended by this merge so far
  unsigned long long t_cmd;               // %globaltimer when the command of this merge arrived
  // per CTA, per merge
  unsigned int n_births, n_recs, n_occ, rec_base, birth_base, n_cand, n_occ1, inbox_n, n_ovf, spec /* this merge was started from a hint */;
  unsigned long long maxpush;          // per CTA, per merge: largest new frequency >= min_freq among this CTA's records (what the host will push)
  // cluster-wide, live in CTA 0 only
  unsigned int spill, n_recs_total, removed, inserted;
  unsigned long long part_cx[CL_SIZE], part_cs[CL_SIZE];  // per-CTA record checksums (plain remote stores; combined by CTA 0)
  unsigned long long part_max[CL_SIZE];                    // per-CTA maxpush
};

struct ClusterSmem {
  unsigned long long *keys, *val, *mk;   // T2: the pairs this CTA owns (owner = pair hash), filled from the inbox
  unsigned long long *k1, *v1, *m1;      // T1: partial sums of the deltas produced by this CTA's own threads
  uint4 *inbox;                          // 2 x uint4 per message: {key, val}, {min key, -}
  unsigned short *occ1;
  uint4 *births;
  Rec *recs;
  unsigned short *occ;                   // occupied slots of this CTA's part
  uint4 *cand;                           // candidate words of this CTA: {word index, header location lo, hi, -}
  ClusterCtl *ctl;
  unsigned long long *csum;              // [64] scratch of block_checksum
  // GRID mode scratch (aliases the delta table, which is empty then)
  int (*rows)[ROW];
  Match (*ml)[MATCH_CAP];
  unsigned int *n_match;
  Rec *tail_stage;
  unsigned int *tail_count;
};
constexpr size_t CL_SMEM_TABLE = (size_t)CL_HT_SLOTS * 24 * 2;
constexpr size_t CL_SMEM_GRID = (size_t)CL_WARPS * ROW * 4 + (size_t)CL_WARPS * MATCH_CAP * sizeof(Match) + CL_WARPS * 4 + STAGE_RECS * sizeof(Rec) + 16;
static_assert(CL_SMEM_GRID <= CL_SMEM_TABLE, "GRID-mode scratch must fit into the delta table's shared memory");
constexpr size_t CL_SMEM_BYTES = CL_SMEM_TABLE + (size_t)CL_BIRTH_STAGE * 16 + (size_t)CL_REC_STAGE * sizeof(Rec) + CL_HT_SLOTS * 4 + (size_t)CL_CAND_CAP * 16 + (size_t)CL_INBOX * 32 +
                                 sizeof(ClusterCtl) + 64 * 8 + 64;

__device__ __forceinline__ ClusterSmem cluster_smem(unsigned char *base) {
  ClusterSmem m;
  unsigned char *p = base;
  m.keys = reinterpret_cast<unsigned long long *>(p);
  m.val = m.keys + CL_HT_SLOTS;
  m.mk = m.val + CL_HT_SLOTS;
  m.k1 = m.mk + CL_HT_SLOTS;
  m.v1 = m.k1 + CL_HT_SLOTS;
  m.m1 = m.v1 + CL_HT_SLOTS;
  {  // GRID-mode view of the same bytes
    unsigned char *q = base;
    m.rows = reinterpret_cast<int (*)[ROW]>(q); q += (size_t)CL_WARPS * ROW * 4;
    m.ml = reinterpret_cast<Match (*)[MATCH_CAP]>(q); q += (size_t)CL_WARPS * MATCH_CAP * sizeof(Match);
    m.tail_stage = reinterpret_cast<Rec *>(q); q += STAGE_RECS * sizeof(Rec);
    m.n_match = reinterpret_cast<unsigned int *>(q); q += CL_WARPS * 4;
    m.tail_count = reinterpret_cast<unsigned int *>(q);
  }
  p += CL_SMEM_TABLE;
  m.births = reinterpret_cast<uint4 *>(p); p += (size_t)CL_BIRTH_STAGE * 16;
  m.recs = reinterpret_cast<Rec *>(p); p += (size_t)CL_REC_STAGE * sizeof(Rec);
  m.cand = reinterpret_cast<uint4 *>(p); p += (size_t)CL_CAND_CAP * 16;
  m.inbox = reinterpret_cast<uint4 *>(p); p += (size_t)CL_INBOX * 32;
  m.csum = reinterpret_cast<unsigned long long *>(p); p += 64 * 8;
  m.ctl = reinterpret_cast<ClusterCtl *>(p); p += sizeof(ClusterCtl);
  m.occ = reinterpret_cast<unsigned short *>(p); p += CL_HT_SLOTS * 2;
  m.occ1 = reinterpret_cast<unsigned short *>(p);
  return m;
}

// (key, delta, first-touch key) -> a table in THIS CTA's shared memory. false: no room within the probe limit.
__device__ __forceinline__ bool smem_table_add(unsigned long long *keys, unsigned long long *val, unsigned long long *mk,
                                               unsigned long long k, long long delta, unsigned long long key,
                                               unsigned short *occ, unsigned int *n_occ, int max_probes,
                                               const PairTableDev::GSlotRef pf = PairTableDev::GSlotRef{nullptr, 0}) {
  uint32_t sl = (uint32_t)(dmix64(k) >> 3) & (CL_HT_SLOTS - 1);
#pragma unroll 1
  for (int probe = 0; probe < max_probes; probe++) {
    const unsigned long long cur = atomicCAS(&keys[sl], PT_EMPTY, k);
    if (cur == PT_EMPTY) occ[atomicAdd(n_occ, 1u)] = (unsigned short)sl;  // the list of occupied slots, for whoever empties the table
    if (cur == PT_EMPTY && pf.slots)  // first touch: pull the pair's frequency-table slot into L2 for the emit phase
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<char *>(pf.slots) + 32ull * ((uint32_t)dmix64(k + 0x632BE59BD9B4E019ull) & pf.mask)));
    if (cur == PT_EMPTY || cur == k) {
      {  // 64-bit add as two native 32-bit shared-memory atomics (a 64-bit atomicAdd on shared memory is a compare-and-swap
         // loop, which crawls when many threads hit one hot pair): every add carries its own overflow into the high half
        unsigned int *v32 = reinterpret_cast<unsigned int *>(&val[sl]);
        const unsigned int dlo = (unsigned int)(unsigned long long)delta, dhi = (unsigned int)((unsigned long long)delta >> 32);
        const unsigned int old = atomicAdd(&v32[0], dlo);
        const unsigned int carry = (old + dlo) < old ? 1u : 0u;
        if (dhi + carry) atomicAdd(&v32[1], dhi + carry);
      }
      unsigned long long sn = *(volatile unsigned long long *)&mk[sl];  // 64-bit min by compare-and-swap
#pragma unroll 1
      while (key < sn) {
        const unsigned long long prev = atomicCAS(&mk[sl], sn, key);
        if (prev == sn) break;
        sn = prev;
      }
      return true;
    }
    sl = (sl + 1) & (CL_HT_SLOTS - 1);
  }
  return false;
}

// deltas -> T1 (this CTA's shared memory: fast local atomics), births -> this CTA's stage
struct ClusterSink {
  cg::cluster_group &cluster;
  const ClusterSmem &m;
  const PairTableDev &t;      // spill target + canonicalisation of a negative unk_id
  const BirthLogDev &lg;
  const ClusterTune &tune;
  __device__ __forceinline__ void add_batch(int n, const int32_t *x, const int32_t *y, const long long *delta, const uint64_t *key) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
      int32_t xx = x[i];
      if (t.canon_on && y[i] == UNK_CODE) xx = t.canon_first;
      const unsigned long long k = ((unsigned long long)(uint32_t)xx << 32) | (uint32_t)y[i];
      if (!smem_table_add(m.k1, m.v1, m.m1, k, delta[i], (unsigned long long)key[i], m.occ1, &m.ctl->n_occ1, (int)tune.max_probes, t.gpf)) {
        // T1 is (nearly) full: this delta goes to the global table, and so will everything else of this merge
        atomicOr(&cluster.map_shared_rank(m.ctl, 0)->spill, 1u);
        pt_add(t, xx, y[i], delta[i], key[i]);
      }
    }
  }
  __device__ __forceinline__ void birth(uint32_t other, bool right_side, uint32_t wi, uint64_t hloc) {
    const uint4 e = log_entry(other, right_side, wi, hloc);
    const unsigned int i = atomicAdd(&m.ctl->n_births, 1u);
    if (i < tune.birth_stage) m.births[i] = e;
    else {  // past the stage: straight to its final place
      ClusterCtl *c0 = cluster.map_shared_rank(m.ctl, 0);
      const unsigned int idx = c0->log_cursor + atomicAdd(&c0->births_total, 1u);
      if (idx < lg.cap) lg.ent[idx] = e;
      else atomicOr(lg.flags, 1u);
    }
  }
};

__device__ __forceinline__ void cluster_barrier(cg::cluster_group &cluster) { cluster.sync(); }
__device__ __forceinline__ void cluster_clear_tables(const ClusterSmem &m) {
  for (int sl = threadIdx.x; sl < CL_HT_SLOTS; sl += CL_THREADS) {
    m.keys[sl] = PT_EMPTY; m.val[sl] = 0; m.mk[sl] = ~0ull;
    m.k1[sl] = PT_EMPTY; m.v1[sl] = 0; m.m1[sl] = ~0ull;
  }
}

// LOCAL merge, exchange: every partial sum of T1 goes to the CTA that owns its pair -- straight into T2 when that is
// this CTA, otherwise as one message into the owner's inbox (one remote 32-bit atomic for the place, two remote
// 16-byte stores). Remote 64-bit min/xor atomics are avoided altogether (wrong results on this toolchain).
__device__ __forceinline__ void cluster_exchange(const ClusterSmem &m, const PairTableDev &t, cg::cluster_group &cluster, unsigned int crank, const ClusterTune &tune) {
  const unsigned int n1 = m.ctl->n_occ1;  // (listed while the deltas were added; the caller has synchronised the block)
#pragma unroll 1
  for (unsigned int i = threadIdx.x; i < n1; i += CL_THREADS) {
    const int sl = m.occ1[i];
    const unsigned long long k = m.k1[sl], d = m.v1[sl], mk = m.m1[sl];
    m.k1[sl] = PT_EMPTY; m.v1[sl] = 0; m.m1[sl] = ~0ull;
    const unsigned int owner = (unsigned int)dmix64(k) & (CL_SIZE - 1);
    bool ok;
    if (owner == crank) ok = smem_table_add(m.keys, m.val, m.mk, k, (long long)d, mk, m.occ, &m.ctl->n_occ, (int)tune.max_probes);
    else {
      ClusterCtl *oc = cluster.map_shared_rank(m.ctl, owner);
      const unsigned int j = atomicAdd(&oc->inbox_n, 1u);
      ok = j < tune.inbox_cap;
      if (ok) {
        uint4 *ob = cluster.map_shared_rank(m.inbox, owner);
        ob[2 * j] = make_uint4((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)d, (uint32_t)(d >> 32));
        ob[2 * j + 1] = make_uint4((uint32_t)mk, (uint32_t)(mk >> 32), 0u, 0u);
      }
    }
    if (!ok) {
      atomicOr(&cluster.map_shared_rank(m.ctl, 0)->spill, 1u);
      pt_add(t, (int32_t)(k >> 32), (int32_t)(k & 0xFFFFFFFFu), (long long)d, mk);
    }
  }
}
// ... and after the cluster barrier the owner folds its inbox into T2
__device__ __forceinline__ void cluster_fold_inbox(const ClusterSmem &m, const PairTableDev &t, cg::cluster_group &cluster, const ClusterTune &tune) {
  const unsigned int nin = min(m.ctl->inbox_n, tune.inbox_cap);
#pragma unroll 1
  for (unsigned int i = threadIdx.x; i < nin; i += CL_THREADS) {
    const uint4 q0 = m.inbox[2 * i], q1 = m.inbox[2 * i + 1];
    const unsigned long long k = ((unsigned long long)q0.y << 32) | q0.x, d = ((unsigned long long)q0.w << 32) | q0.z,
                             mk = ((unsigned long long)q1.y << 32) | q1.x;
    if (!smem_table_add(m.keys, m.val, m.mk, k, (long long)d, mk, m.occ, &m.ctl->n_occ, CL_MAX_PROBES * 8)) {
      atomicOr(t.flags, 1u);  // cannot happen (T2 is at most 2/3 full with a whole inbox): reported as an internal sizing error
    }
  }
  __syncthreads();
}

// LOCAL merge, phase 2 of one CTA: apply this CTA's pairs to the device frequency table and stage the records
__device__ __forceinline__ void cluster_emit_part(const ClusterSmem &m, const EmitMode &em, const PairTableDev &t, bool spill,
                                                  unsigned long long &cx, unsigned long long &cs, unsigned int &inserted,
                                                  unsigned long long &maxpush, Rec *__restrict__ out, size_t out_cap, cg::cluster_group &cluster, int32_t new_id, const ClusterTune &tune,
                                                  unsigned int n_occ /* entries listed in m.occ (the caller has synchronised the block) */) {
  ClusterCtl *c0 = cluster.map_shared_rank(m.ctl, 0);
#pragma unroll 1
  for (unsigned int i = threadIdx.x; i < n_occ; i += CL_THREADS) {
    const int sl = m.occ[i];
    const unsigned long long k = m.keys[sl];
    const long long d = (long long)m.val[sl];
    const unsigned long long mk = m.mk[sl];
    m.keys[sl] = PT_EMPTY; m.val[sl] = 0; m.mk[sl] = ~0ull;
    if (spill) {  // some delta of this merge went to the global table: everything follows it there
      pt_add(t, (int32_t)(k >> 32), (int32_t)(k & 0xFFFFFFFFu), d, mk);
      continue;
    }
    if (k == em.merged_key) continue;  // reference bpe.cpp:494-496
    uint32_t g = gt_home(em.g, k);
    unsigned long long old;
    if ((int32_t)(k >> 32) == new_id || (int32_t)(k & 0xFFFFFFFFu) == new_id) {
      // a pair around the token this merge creates cannot be in the table yet: claim its slot straight away
      const unsigned int before = inserted;
      g = gt_upsert(em.g, k, em.stamp_base | delta_bucket(em, k), ~mk, inserted);
      old = inserted != before ? 0ull : em.g.slots[g].freq;
    } else {
      const ulonglong2 gs = __ldcg(reinterpret_cast<const ulonglong2 *>(em.g.slots + g));
      if (gs.x == k) old = gs.y;
      else {
        g = gt_upsert(em.g, k, em.stamp_base | delta_bucket(em, k), ~mk, inserted);
        old = em.g.slots[g].freq;
      }
    }
    unsigned long long nw;
    if (d < 0) { const unsigned long long ad = (unsigned long long)(-d); nw = old >= ad ? old - ad : 0ull; }  // bpe.cpp:500-509
    else nw = old + (unsigned long long)d;
    em.g.slots[g].freq = nw;
    if (old >= em.min_freq || nw >= em.min_freq) {
      if (nw >= em.min_freq && nw > maxpush) maxpush = nw;  // the host pushes this pair (reference bpe.cpp:512-515)
      const unsigned int j = atomicAdd(&m.ctl->n_recs, 1u);
      if (j < tune.rec_stage) rec_out(m.recs, CL_REC_STAGE, j, k, (long long)nw, mk, cx, cs);
      else rec_out(out, out_cap, atomicAdd(&c0->n_recs_total, 1u), k, (long long)nw, mk, cx, cs);  // past the stage: straight to its final place
    }
  }
}

__global__ void __launch_bounds__(CL_THREADS, 1)
merge_cluster(StreamDev s, PairTableDev t, EmitMode em, unsigned long long *removed_total, Rec *__restrict__ out0, Rec *__restrict__ out1, size_t out_cap,
              unsigned long long *__restrict__ out_hdr0, unsigned long long *__restrict__ out_hdr1, unsigned long long seq_base, unsigned long long op_base,
              volatile HostCmd2 *hcmd /* [8] in mapped host memory: [0] the command, [1], [2] the hints for even / odd sequence numbers (host -> device); [4] status (device -> host) */, DevCmd2 *dcmd, unsigned long long timeout_ns, unsigned long long *trace,
              uint4 *ovf /* [CL_LOCAL_MAX]: candidate words that did not fit a CTA's shared-memory list */,
              const HostCmd2 *__restrict__ script /* nullptr, or [script_n] commands in DEVICE memory that replace the mailbox: the kernel then runs without the host (profiling under ncu's kernel replay, see TrainerImpl::profile_scripted) */, unsigned long long script_n, ClusterTune tune,
              unsigned long long *acct /* [32]: LOCAL merges, their device ns, GRID merges, their device ns, hints accepted, hints rejected, hints accepted without a PCIe trip, LOCAL merges that spilled to the global table, [8] sequence number the watchdog gave up on, [9] after how many ns */) {
  extern __shared__ __align__(16) unsigned char cl_dyn_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const ClusterSmem m = cluster_smem(cl_dyn_smem);
  __shared__ bool is_last;
  const unsigned int crank = cluster.block_rank();
  const bool leader = blockIdx.x < CL_SIZE;  // cluster 0
  const int lane = threadIdx.x & 31;
  const uint32_t log_m_base = em.log.m_cur;

  cluster_clear_tables(m);
  if (threadIdx.x == 0) memset(m.ctl, 0, sizeof(ClusterCtl));
  __syncthreads();
  cluster_barrier(cluster);

  unsigned long long grid_epoch = 0;  // GRID merges seen so far
  unsigned int pre_flags = 0;
  bool cursor_stale = true;  // (CTA 0 / thread 0) the log cursor kept in shared memory is behind the global one
  // (CTA 0 / thread 0) hints: the previous merge was LOCAL and clean, the largest frequency it makes the host push, its token
  bool hint_ok = false;
  unsigned long long prev_maxpush = 0;
  unsigned int prev_new_id = 0;
  bool grid_pending = false;  // (CTA 0 / thread 0) the previous merge was GRID: its tail (any block) reports through dcmd->tail_max / tail_flags
  unsigned int grid_new_id = 0;
  unsigned int last_pub_lo = 0, last_pub_path = 0;  // (CTA 0 / thread 0) low half of the last sequence number it published itself, and through which path (1 LOCAL/SOLO, 2 spill tail)
  unsigned int go_sent = 0, go_seen = 0;  // (CTA 0 / thread 0) commands handed to the other CTAs of the leader cluster; (other CTAs / thread 0) seen
  uint4 pre_hv = make_uint4(0u, 0u, 0u, 0u);  // the hint word for the NEXT merge, requested while this one runs (a read of mapped host memory takes microseconds)
  long long tr_poll = 0, tr_p1 = 0, tr_p2 = 0, tr_pub = 0;
  for (unsigned long long k = 0;; k++) {
    // ---------------------------------------------------------------- next command
    if (leader) {
      if (crank == 0 && threadIdx.x == 0) {
        const long long c0 = clock64();
        const unsigned long long want = seq_base + k + 1;
        const unsigned long long t0 = gtime_ns();
        unsigned long long pair = 0, nio = 3ull << 32, lr = ~0ull;
        unsigned int spec = 0;
        if (grid_pending && script == nullptr) {
          // The merge before was GRID: whichever block finished last ran its tail. The host answers only once it has seen that
          // tail's header, so waiting for the tail costs nothing -- and with what the tail reports (largest pushed frequency,
          // flags) the merge after a GRID merge can start from a hint like any other.
          while (*(volatile unsigned long long *)&dcmd->tail_done < grid_epoch) { if (gtime_ns() - t0 > timeout_ns) break; }
          __threadfence();
          if (*(volatile unsigned long long *)&dcmd->tail_done >= grid_epoch) {
            const unsigned long long fl = *(volatile unsigned long long *)&dcmd->tail_flags;
            hint_ok = (fl & 0xFFFFFFFFull & ~64ull) == 0ull;
            prev_maxpush = *(volatile unsigned long long *)&dcmd->tail_max;
            prev_new_id = grid_new_id;
          }
          grid_pending = false;
        }
        for (unsigned long long spin = 0;; spin++) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u), hv = make_uint4(0u, 0u, 0u, 0u);  // one 16-byte load from mapped host memory = one PCIe read (the two are in flight together)
          bool have_v = true;
          if (spin == 0 && hint_ok && pre_hv.z != 0u && pre_hv.w == cmd3_word(want, pre_hv.x, pre_hv.y, pre_hv.z)) { hv = pre_hv; have_v = false; }  // already here: no trip at all
          else if (script) {
            // (a host only sends the next command once it has seen the result; a script has to wait itself until the tail of
            //  the GRID merge before -- run by whichever block finished last, possibly in another cluster -- is through)
            while (*(volatile unsigned long long *)&dcmd->tail_done < grid_epoch) { if (gtime_ns() - t0 > timeout_ns) break; }
            __threadfence();
            if (k >= script_n) { nio = 1ull << 32; break; }  // end of the script: stop
            v = __ldcg(reinterpret_cast<const uint4 *>(script + k));
          } else {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(hcmd) : "memory");
            if (hint_ok) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w) : "l"(hcmd + 1 + (want & 1ull)) : "memory");
          }
          if (have_v && v.w == cmd3_word(want, v.x, v.y, v.z)) {
            pair = ((unsigned long long)v.y << 32) | v.x;
            nio = (unsigned long long)(v.z & 0x0FFFFFFFu) | ((unsigned long long)(v.z >> 28) << 32);
            break;
          }
          if (have_v && (v.z >> 28) == 1u && v.w == cmd3_word(want - 1, v.x, v.y, v.z)) {  // a stop addressed to the merge this kernel started from a hint
            nio = 1ull << 32;
            break;
          }
          // Hint {second, first, frequency F}: "the next pair is (first, second) if the merge before it pushes nothing >= F and
          // leaves that pair's frequency at F" (the host derived it from its exact heap, see HostCore::peek_next). Both conditions
          // are checked here, against the device frequency table; a rejected hint is ignored and the command awaited.
          if (hint_ok && hv.z != 0u && hv.w == cmd3_word(want, hv.x, hv.y, hv.z)) {
            const unsigned long long hk = ((unsigned long long)hv.y << 32) | hv.x, F = hv.z;
            if (prev_maxpush < F && gt_find_freq(em.g, hk) == F) {
              pair = hk; nio = (unsigned long long)((prev_new_id + 1u) & 0x0FFFFFFFu); spec = 1u;
              acct[4] += 1;
              if (!have_v) acct[6] += 1;  // ... and it was already here when the merge before finished
              break;
            }
            acct[5] += 1;
            hint_ok = false;
          }
          if ((spin & 63) == 63 && gtime_ns() - t0 > timeout_ns) { acct[8] = want; acct[9] = gtime_ns() - t0; break; }  // abort: the host went away (the host reads why)
        }
        hint_ok = false;
        if (!(nio >> 32)) {  // the birth log of the newer token of the pair (device-side bookkeeping: cheaper than a second PCIe trip)
          const int32_t pa = (int32_t)(pair >> 32), pb = (int32_t)(pair & 0xFFFFFFFFu), newer = pa > pb ? pa : pb;
          const uint32_t mcur = (uint32_t)((uint32_t)(nio & 0xFFFFFFFFu) - 256u);
          if (em.log.ent != nullptr && newer >= 256 && (uint32_t)(newer - 256) < mcur) {
            const unsigned int lo = __ldcg(&em.log.start[newer - 256]), hi = __ldcg(&em.log.start[newer - 256 + 1]);
            if (hi >= lo && hi <= em.log.cap) lr = ((unsigned long long)lo << 32) | (unsigned long long)(hi - lo);
            else atomicOr(em.log.flags, 2u);  // a log range that cannot be: reported to the host (header flag 32) instead of followed; this merge scans the rows
          } else if (em.log.ent != nullptr && ip_lookup(em.log, pa, pb)) {  // two initial symbols: their occurrence index, unless the list is long (then the whole grid scans)
            const unsigned int pk = (unsigned int)pa * 256u + (unsigned int)pb;
            const unsigned int lo = __ldcg(&em.log.ip_start[pk]), hi = __ldcg(&em.log.ip_start[pk + 1]);
            if (hi - lo <= em.log.ip_local_max) lr = ((unsigned long long)lo << 32) | (unsigned long long)(hi - lo);
          }
          if (cursor_stale && em.log.ent != nullptr) { m.ctl->log_cursor = __ldcg(em.log.cursor); cursor_stale = false; }
          m.ctl->births_total = 0;
        }
        const unsigned long long t_cmd = gtime_ns();
        const unsigned int stop = (unsigned int)(nio >> 32);
        const unsigned int ln = (unsigned int)(lr & 0xFFFFFFFFu);
        const unsigned int mode = (lr == ~0ull || ln > tune.local_max) ? 1u : (!stop && ln <= tune.solo_max) ? 2u : 0u;
        // SOLO: only this CTA learns about the merge; the other CTAs of the cluster keep waiting for their `go`
        for (unsigned int r = 0; r < (mode == 2u ? 1u : (unsigned int)CL_SIZE); r++) {
          ClusterCtl *c = cluster.map_shared_rank(m.ctl, r);
          c->pair = pair; c->new_id_op = nio; c->log_range = lr; c->k = k; c->mode = mode; c->stop = stop; c->t_cmd = t_cmd; c->spec = spec;
        }
        if (mode != 2u) {
          asm volatile("fence.acq_rel.cluster;" ::: "memory");
          go_sent++;
          for (unsigned int r = 1; r < CL_SIZE; r++) *(volatile unsigned int *)&cluster.map_shared_rank(m.ctl, r)->go = go_sent;
        }
        if (mode == 1u) cursor_stale = true;  // a GRID merge appends through the global cursor
        if (mode == 1u && !stop) { grid_pending = true; grid_new_id = (unsigned int)(nio & 0x0FFFFFFFull); }
        if (stop || mode == 1u) {  // the other clusters take part (or leave)
          dcmd->tail_max = 0ull;
          dcmd->pair = pair; dcmd->new_id_op = nio; dcmd->log_range = lr; dcmd->k = k; dcmd->t_cmd = t_cmd; dcmd->spec = spec;
          __threadfence();
          *(volatile unsigned long long *)&dcmd->epoch = grid_epoch + 1;
        }
        *(volatile unsigned long long *)&dcmd->alive_ns = gtime_ns();
        {  // device -> host status (mailbox word 4, its own 64-byte line): which command was just accepted and how, what was
           // published last. Diagnostics only: the host prints it when a result does not arrive.
          const unsigned int sx = (unsigned int)want, sy = mode | (spec << 8) | (stop << 16) | ((unsigned int)(grid_epoch & 0xFFFFu) << 20), sz = last_pub_lo, sw = last_pub_path;
          asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(hcmd + 4), "r"(sx), "r"(sy), "r"(sz), "r"(sw) : "memory");
        }
        tr_poll += clock64() - c0;
        // the hint for the merge after this one may already be in the mailbox (the host sends it a merge ahead when it can):
        // ask for it now, look at it when this merge is done
        if (!stop) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pre_hv.x), "=r"(pre_hv.y), "=r"(pre_hv.z), "=r"(pre_hv.w) : "l"(hcmd + 1 + ((want + 1) & 1ull)) : "memory");
      }
      if (crank != 0 && threadIdx.x == 0) {  // wait for a command that concerns the whole cluster (SOLO merges pass these CTAs by)
        while (*(volatile unsigned int *)&m.ctl->go == go_seen) { }
        go_seen = *(volatile unsigned int *)&m.ctl->go;
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
      }
      __syncthreads();
    } else {
      if (threadIdx.x == 0) {
        unsigned long long pair = 0, nio = 3ull << 32, lr = ~0ull, kk = 0;
        for (unsigned long long spin = 0;; spin++) {
          if (*(volatile unsigned long long *)&dcmd->epoch > grid_epoch) {
            __threadfence();
            pair = *(volatile unsigned long long *)&dcmd->pair; nio = *(volatile unsigned long long *)&dcmd->new_id_op;
            lr = *(volatile unsigned long long *)&dcmd->log_range; kk = *(volatile unsigned long long *)&dcmd->k;
            m.ctl->t_cmd = *(volatile unsigned long long *)&dcmd->t_cmd;
            m.ctl->spec = (unsigned int)*(volatile unsigned long long *)&dcmd->spec;
            break;
          }
          if ((spin & 255) == 255) {  // leader gone?
            // %globaltimer of two SMs can differ by a tick: the leader's stamp may be AHEAD of this SM's clock, and an unsigned
            // difference would then be huge -- this block would leave, the next GRID merge would wait for it for ever (144
            // blocks expected, 143 arrive, nobody is last, nothing is published). Signed difference; 0 = the leader has not
            // accepted its first command yet.
            const unsigned long long alive = *(volatile unsigned long long *)&dcmd->alive_ns;
            if (alive != 0ull && (long long)(gtime_ns() - alive) > (long long)(4 * timeout_ns)) break;
          }
        }
        m.ctl->pair = pair; m.ctl->new_id_op = nio; m.ctl->log_range = lr; m.ctl->k = kk; m.ctl->mode = 1u; m.ctl->stop = (unsigned int)(nio >> 32);
      }
      __syncthreads();
    }
    if (m.ctl->stop) return;
    const unsigned long long pairk = m.ctl->pair;
    const int32_t a = (int32_t)(pairk >> 32), b = (int32_t)(pairk & 0xFFFFFFFFu), new_id = (int32_t)(m.ctl->new_id_op & 0xFFFFFFFFu);
    const unsigned int mode = m.ctl->mode;
    em.log.m_cur = (uint32_t)(new_id - 256);  // (== log_m_base + merges done; the host only sends consistent ids while the log is on)
    em.merged_key = pairk;
    const unsigned long long mk_ = m.ctl->k;  // (the other clusters only iterate on GRID merges: their own counter lags)
    em.stamp_base = (op_base + mk_) << 10;
    const unsigned long long seq = seq_base + mk_ + 1;
    // two record buffers / headers, used in turn: a merge started from a hint writes its results while the host still reads
    // those of the merge before it
    Rec *__restrict__ out = (seq & 1ull) ? out1 : out0;
    unsigned long long *__restrict__ out_hdr = (seq & 1ull) ? out_hdr1 : out_hdr0;
    const unsigned int spec_flag = m.ctl->spec ? 64u : 0u;

    if (mode == 1u) {
      // ---------------------------------------------------------------- GRID: all clusters, global pair table
      grid_epoch++;
      uint32_t removed = scan_merge(s, t, em.log, a, b, new_id, m.rows, m.ml, m.n_match);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) removed += __shfl_down_sync(0xffffffffu, removed, d);
      if (lane == 0 && removed) atomicAdd(removed_total, (unsigned long long)removed);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(&dcmd->done_count, 1ull) + 1ull == grid_epoch * (unsigned long long)gridDim.x;
      }
      __syncthreads();
      if (is_last) {
        __threadfence();
        TailSmem ts{m.tail_stage, m.csum, m.tail_count};
        fused_tail(t, em, ts, out, out_cap, out_hdr, removed_total, seq, spec_flag, nullptr, &dcmd->tail_max, &dcmd->tail_flags);
        if (threadIdx.x == 0) {  // (one publisher at a time)
          __threadfence();
          *(volatile unsigned long long *)&dcmd->tail_done = grid_epoch;
          acct[10] = seq; acct[11] = blockIdx.x;  // (diagnostics: the last GRID merge whose tail ran, and where)
          const long long dts = (long long)(gtime_ns() - m.ctl->t_cmd);  // (stamps of two SMs: may come out a tick negative)
          const unsigned long long dt = dts > 0 ? (unsigned long long)dts : 0ull;
          acct[2] += 1; acct[3] += dt;
          if (trace) {  // development aid: how long do GRID merges take? [16..23] = counts, [24..31] = ns, by duration class
            const int cls = dt < 16000 ? 0 : dt < 24000 ? 1 : dt < 32000 ? 2 : dt < 48000 ? 3 : dt < 64000 ? 4 : dt < 128000 ? 5 : dt < 256000 ? 6 : 7;
            trace[16 + cls] += 1; trace[24 + cls] += dt;
          }
        }
      }
      __syncthreads();
      // the scratch aliased the delta tables: empty them again
      cluster_clear_tables(m);
      __syncthreads();
      continue;
    }

    // ---------------------------------------------------------------- LOCAL: the leader cluster alone; SOLO: its CTA 0 alone
    // (SOLO = the same code with one CTA's worth of log entries, the deltas emitted straight from this CTA's own table T1,
    //  no exchange through the other CTAs' shared memory and no cluster barrier: most merges touch a few hundred words)
    const bool solo = mode == 2u;
    const unsigned int n_cta = solo ? 1u : (unsigned int)CL_SIZE, my_cta = solo ? 0u : crank;
    const long long c1 = clock64();
    {
      const unsigned long long lr = m.ctl->log_range;
      const uint64_t lo = lr >> 32, n = lr & 0xFFFFFFFFu;
      uint32_t merge = 0, other = 0, side = 0;
      bool have_log = log_lookup(em.log, a, b, merge, other, side);  // (true: the command came with a log range)
      const uint4 *__restrict__ ent = em.log.ent;
      if (!have_log && ip_lookup(em.log, a, b)) { have_log = true; other = (uint32_t)a; side = 0u; ent = em.log.ip_ent; }  // the occurrence index has the same format
      ClusterSink sink{cluster, m, t, em.log, tune};
      uint32_t removed = 0;
      // Eight log entries per thread are requested together (one round trip); the matching ones are listed in
      // shared memory and then dealt out evenly, so that no thread rewrites more words than its neighbours.
      long long cA = 0;
      for (uint64_t base = 0; have_log && base < n; base += 8ull * n_cta * CL_THREADS) {
        uint4 ev[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint64_t i = base + (uint64_t)u * (n_cta * CL_THREADS) + my_cta * CL_THREADS + threadIdx.x;
          ev[u] = (i < n && SWB_DBG_OK(ent == em.log.ip_ent || lo + i < em.log.cap, 1, lo, i, n)) ? __ldcg(&ent[lo + i]) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint4 e = ev[u];
          if (e.x == other && (e.y & 0x80000000u) == side && !(e.x == 0xFFFFFFFFu && e.y == 0xFFFFFFFFu)) {
            const unsigned int ci = atomicAdd(&m.ctl->n_cand, 1u);
            {  // the word and its count are on their way while the candidates are dealt out
              const uint64_t hl = ((uint64_t)e.w << 32) | e.z;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const int32_t *>(s.rows) + hl));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(s.cnt + (e.y & 0x7FFFFFFFu)));
            }
            if (ci < tune.cand_cap) m.cand[ci] = make_uint4(e.y & 0x7FFFFFFFu, e.z, e.w, 0u);
            else {  // rare: listed in global memory
              const unsigned int oi = atomicAdd(&m.ctl->n_ovf, 1u);
              if (SWB_DBG_OK(oi < CL_LOCAL_MAX / CL_SIZE, 2, oi, ci, n)) ovf[(size_t)crank * (CL_LOCAL_MAX / CL_SIZE) + oi] = make_uint4(e.y & 0x7FFFFFFFu, e.z, e.w, 0u);
            }
          }
        }
      }
      __syncthreads();
      if (trace && crank == 0 && threadIdx.x == 0) { cA = clock64(); trace[8] += (unsigned long long)(cA - c1); }
      {
        // candidate i goes to warp i % 16, lane i / 16: the work is spread over all warps (and schedulers) of the CTA;
        // candidates past the shared-memory list (rare) follow from this CTA's part of the global overflow list
        const unsigned int nc = min(m.ctl->n_cand, tune.cand_cap), nall = nc + m.ctl->n_ovf;
        const uint4 *ovf_mine = ovf + (size_t)crank * (CL_LOCAL_MAX / CL_SIZE);
#pragma unroll 1
        for (unsigned int i = (threadIdx.x >> 5) + (unsigned int)CL_WARPS * (threadIdx.x & 31); i < nall; i += CL_THREADS) {
          const uint4 cnd = i < nc ? m.cand[i] : __ldcg(&ovf_mine[i - nc]);
          if (SWB_DBG_OK((((uint64_t)cnd.z << 32) | cnd.y) < s.n_rows * ROW, 3, ((uint64_t)cnd.z << 32) | cnd.y, cnd.x, i))
            removed += merge_one_word(s, ((uint64_t)cnd.z << 32) | cnd.y, cnd.x, a, b, new_id, sink);
        }
      }
      if (trace && crank == 0 && threadIdx.x == 0) trace[9] += (unsigned long long)(clock64() - cA);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) removed += __shfl_down_sync(0xffffffffu, removed, d);
      if (lane == 0 && removed) atomicAdd(&cluster.map_shared_rank(m.ctl, 0)->removed, removed);
      if (trace && crank == 0 && threadIdx.x == 0) { trace[5] += (unsigned long long)(clock64() - c1); trace[6] += n; }
    }
    __syncthreads();
    if (trace && crank == 0 && threadIdx.x == 0) trace[7] += (unsigned long long)(clock64() - c1);
    if (!solo) {
      cluster_exchange(m, t, cluster, crank, tune);
      cluster_barrier(cluster);  // every partial sum of this merge has reached the CTA that owns its pair
    }
    const long long c2 = clock64();
    {
      ClusterCtl *c0 = cluster.map_shared_rank(m.ctl, 0);
      const bool spill = c0->spill != 0;
      if (crank == 0 && threadIdx.x == 0) {
        // rarely set conditions, requested now and looked at when the header is written (a flag raised later in this
        // merge is seen one merge later: the thresholds leave that much room, the others are fatal either way)
        pre_flags = (__ldcg(t.flags) & 1u) | (__ldcg(em.log.flags) ? 2u : 0u) |
                    ((__ldcg(em.g.flags) || __ldcg(em.g.n_used) >= (em.g.mask >> 1)) ? 4u : 0u);
        if (pre_hv.z != 0u && pre_hv.w == cmd3_word(seq + 1, pre_hv.x, pre_hv.y, pre_hv.z))  // the hinted pair's slot: looked at when this merge is done
          gt_prefetch(em.g, ((unsigned long long)pre_hv.y << 32) | pre_hv.x);
      }
      if (crank == 0 && threadIdx.x == 32 && !spill) {  // the merged pair's frequency becomes 0 (bpe.cpp:523)
        unsigned int ins = 0;
        em.g.slots[gt_upsert(em.g, em.merged_key, em.stamp_base | delta_bucket(em, em.merged_key), 0ull, ins)].freq = 0;
        if (ins) atomicAdd(em.g.n_used, ins);
      }
      if (!solo) cluster_fold_inbox(m, t, cluster, tune);
      unsigned long long cx = 0, cs = 0;
      unsigned int inserted = 0;
      if (threadIdx.x == 64) {  // this CTA's range in the birth log: requested now, needed after the records are staged
        const unsigned int nb = min(m.ctl->n_births, tune.birth_stage);
        m.ctl->birth_base = c0->log_cursor + (nb ? atomicAdd(&c0->births_total, nb) : 0u);
      }
      unsigned long long maxpush = 0;
      if (!solo) cluster_emit_part(m, em, t, spill, cx, cs, inserted, maxpush, out, out_cap, cluster, new_id, tune, m.ctl->n_occ);
      else {  // T1 is the whole delta table of this merge
        ClusterSmem m1 = m;
        m1.keys = m.k1; m1.val = m.v1; m1.mk = m.m1; m1.occ = m.occ1;
        cluster_emit_part(m1, em, t, spill, cx, cs, inserted, maxpush, out, out_cap, cluster, new_id, tune, m.ctl->n_occ1);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        inserted += __shfl_down_sync(0xffffffffu, inserted, d);
        const unsigned long long o = __shfl_down_sync(0xffffffffu, maxpush, d);
        maxpush = o > maxpush ? o : maxpush;
      }
      if (lane == 0 && inserted) atomicAdd(em.g.n_used, inserted);  // (no result needed here: CTA 0 looks at the load once per merge)
      if (lane == 0 && maxpush) atomicMax(&m.ctl->maxpush, maxpush);
      __syncthreads();
      if (threadIdx.x == 0) {  // this CTA's ranges in the record buffer and in the birth log
        const unsigned int nr = min(m.ctl->n_recs, tune.rec_stage);
        m.ctl->rec_base = nr ? atomicAdd(&c0->n_recs_total, nr) : 0u;
      }
      __syncthreads();
      {
        const unsigned int nr = min(m.ctl->n_recs, tune.rec_stage), base = m.ctl->rec_base;
        const uint4 *src = reinterpret_cast<const uint4 *>(m.recs);
        for (unsigned int i = threadIdx.x; i < 2u * nr; i += CL_THREADS) {
          const size_t r = (size_t)base + (i >> 1);
          if (r < out_cap) reinterpret_cast<uint4 *>(out)[2 * r + (i & 1)] = src[i];
        }
        const unsigned int nb = min(m.ctl->n_births, tune.birth_stage), bbase = m.ctl->birth_base;
        for (unsigned int i = threadIdx.x; i < nb; i += CL_THREADS) {
          if (bbase + i < em.log.cap) em.log.ent[bbase + i] = m.births[i];
          else atomicOr(em.log.flags, 1u);
        }
      }
      block_checksum(cx, cs, m.csum);
      if (threadIdx.x == 0) {
        c0->part_cx[crank] = cx; c0->part_cs[crank] = cs; c0->part_max[crank] = m.ctl->maxpush;
        m.ctl->maxpush = 0;
        m.ctl->n_births = 0; m.ctl->n_recs = 0; m.ctl->n_occ = 0; m.ctl->n_cand = 0; m.ctl->n_occ1 = 0; m.ctl->inbox_n = 0; m.ctl->n_ovf = 0;
      }
    }
    if (!solo) cluster_barrier(cluster);  // all records, checksums and log entries of this merge are out
    else __syncthreads();
    const long long c3 = clock64();
    if (crank == 0) {
      ClusterCtl *c = m.ctl;
      if (c->spill) {  // rare: the merge did not fit the distributed table; CTA 0 finishes it from the global one
        if (threadIdx.x == 0) {
          if (c->removed) atomicAdd(removed_total, (unsigned long long)c->removed);
          *em.log.cursor = c->log_cursor + c->births_total;  // (this merge's births were placed through the shared counter)
        }
        __threadfence();
        __syncthreads();
        TailSmem ts{m.tail_stage, m.csum, m.tail_count};
        fused_tail(t, em, ts, out, out_cap, out_hdr, removed_total, seq, spec_flag);
        __syncthreads();
        cluster_clear_tables(m);
        if (threadIdx.x == 0) { acct[0] += 1; acct[1] += gtime_ns() - c->t_cmd; acct[7] += 1; last_pub_lo = (unsigned int)seq; last_pub_path = 2u; }
        if (threadIdx.x == 0) {
          c->log_cursor += c->births_total;  // (the births of this merge are in the log: the next LOCAL merge appends after them)
          c->spill = 0; c->n_recs_total = 0; c->removed = 0;
        }
        __syncthreads();
      } else if (threadIdx.x == 0) {
        const unsigned int n = c->n_recs_total;
        unsigned long long flags = (n > out_cap ? 4u : 0u) | (pre_flags & 1u) | ((pre_flags & 4u) ? 16u : 0u) | spec_flag;
        const unsigned int cur = c->log_cursor + c->births_total;  // (GRID merges append through the global cursor: keep it current)
        *em.log.cursor = cur;
        c->log_cursor = cur;
        em.log.start[em.log.m_cur + 1] = cur;
        if (pre_flags & 2u) flags |= 32u;
        flags |= (unsigned long long)cur << 32;  // the host keeps the log ranges
        unsigned long long x = 0, sm = 0;
        unsigned long long mxp = 0;
#pragma unroll
        for (int r = 0; r < CL_SIZE; r++) {
          if (solo && r > 0) break;  // (the other CTAs' parts are those of an earlier merge)
          x ^= c->part_cx[r]; sm += c->part_cs[r]; mxp = c->part_max[r] > mxp ? c->part_max[r] : mxp;
        }
        // the next merge may start from a hint: this one was LOCAL, complete and raised no flag
        hint_ok = pre_flags == 0u && n <= out_cap && script == nullptr;
        prev_maxpush = mxp; prev_new_id = (unsigned int)new_id;
        {  // the 64-byte header as four 16-byte stores (fewer PCIe writes than eight 8-byte ones; it validates itself)
          const unsigned long long rem = c->removed, chk = hdr_check(seq, n, flags, rem, x, sm);
          asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(out_hdr + 2), "l"(flags), "l"(rem) : "memory");
          asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(out_hdr + 4), "l"(x), "l"(sm) : "memory");
          asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(out_hdr + 0), "l"(seq), "l"((unsigned long long)n) : "memory");
          asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(out_hdr + 6), "l"(chk), "l"(seq) : "memory");
          const unsigned long long dt = gtime_ns() - c->t_cmd;
          acct[0] += 1; acct[1] += dt;
          last_pub_lo = (unsigned int)seq; last_pub_path = 1u;
          {  // LOCAL merges by length of the log they read: [16 + cls] merges, [24 + cls] their ns (cls: <=512, <=4096, <=32768, more entries)
            const unsigned int ln = (unsigned int)(c->log_range & 0xFFFFFFFFu);
            const int cls = ln <= 512u ? 0 : ln <= 4096u ? 1 : ln <= 32768u ? 2 : 3;
            acct[16 + cls] += 1; acct[24 + cls] += dt;
            acct[20 + cls] += n;  // records sent to the host
          }
        }
        c->n_recs_total = 0; c->removed = 0;
        if (trace) {
          trace[0] += 1; trace[1] += (unsigned long long)(c2 - c1); trace[2] += (unsigned long long)(c3 - c2);
          trace[3] += (unsigned long long)(clock64() - c3); trace[4] += (unsigned long long)tr_poll; tr_poll = 0;
        }
      }
    }
    (void)tr_p1; (void)tr_p2; (void)tr_pub; (void)log_m_base;
  }
}

}  // namespace swb
