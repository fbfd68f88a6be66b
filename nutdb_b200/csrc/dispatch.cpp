// dispatch.cpp -- the multi-GPU batch dispatcher behind nutdb_gpu_mctx_* (include/nutdb_gpu.h).
//
// Statements are independent (Parser::parse keeps no state across calls, reference src/parser/mod.rs:21-37), so a
// batch shards by contiguous statement ranges.  A dispatcher owns `workers` contexts on every device it was given, one
// host thread per context.  A call hands it SHARDS (statement ranges with their text); every device's workers take that
// device's shards in order: parse on the device (nutdb_gpu_parse_batch, outputs left in HBM), then move the statement
// records, wire nodes and error records to the GATHER POINT -- a pinned host slot of the worker (one PCIe copy per
// array, straight from the device that produced it) or a slot in device 0's memory (cudaMemcpyPeerAsync: NVLink) -- and
// call the consumer with views of the slot.  While one worker copies down, another one's kernels run and a third's text
// goes up.  No collective: nothing is reduced, results are only gathered.  Indices inside a chunk stay chunk-local (u32);
// `first_stmt` places the chunk in the batch, so nothing overflows however many GPUs share a batch.
//
// Host code only: no kernels here, no parsing on the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nutdb_gpu.h"

namespace {

struct Slot {  // gather buffers of one worker, grow-only
  void* p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // stmt, pnode, err, tok_type, tok_start, tok_end, tok_kw
  size_t cap[7] = {0, 0, 0, 0, 0, 0, 0};
  bool on_device = false;
};

struct Worker {
  int device_index = 0;  // position in the dispatcher's device list
  int device = 0;        // CUDA ordinal
  NutdbCtx* ctx = nullptr;
  Slot host, dev0;
};

}  // namespace

struct NutdbMCtx {
  std::vector<int> devices;
  int workers_per_device = 0;
  std::vector<Worker> workers;
  std::string err;
  std::mutex mu;
  bool peer_checked = false;
};

namespace {

int slot_reserve(NutdbMCtx* m, Slot& s, int i, size_t bytes, bool on_device0) {
  if (bytes <= s.cap[i]) return NUTDB_OK;
  size_t want = bytes + bytes / 8 + 4096;
  if (s.p[i]) {
    if (s.on_device) cudaFree(s.p[i]);
    else cudaFreeHost(s.p[i]);
    s.p[i] = nullptr;
    s.cap[i] = 0;
  }
  cudaError_t e;
  if (on_device0) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(m->devices[0]);
    e = cudaMalloc(&s.p[i], want);
    cudaSetDevice(cur);
  } else {
    e = cudaHostAlloc(&s.p[i], want, cudaHostAllocPortable);
  }
  if (e != cudaSuccess) {
    std::lock_guard<std::mutex> g(m->mu);
    m->err = std::string("gather slot allocation failed: ") + cudaGetErrorString(e);
    s.p[i] = nullptr;
    return NUTDB_E_NOMEM;
  }
  s.cap[i] = want;
  s.on_device = on_device0;
  return NUTDB_OK;
}

void slot_free(NutdbMCtx* m, Slot& s) {
  for (int i = 0; i < 7; i++)
    if (s.p[i]) {
      if (s.on_device) {
        cudaSetDevice(m->devices[0]);
        cudaFree(s.p[i]);
      } else {
        cudaFreeHost(s.p[i]);
      }
      s.p[i] = nullptr;
      s.cap[i] = 0;
    }
}

struct Run {
  NutdbMCtx* m;
  const NutdbMShard* shards;
  uint64_t n_shards;
  uint32_t flags;
  nutdb_chunk_fn fn;
  void* user;
  std::vector<std::vector<uint64_t>> per_device;  // shard indices of each device, in order
  std::vector<std::atomic<uint64_t>> next;        // next position in per_device[d]
  std::atomic<int> rc{NUTDB_OK};
  std::mutex cb_mu;
  explicit Run(size_t ndev) : per_device(ndev), next(ndev) {
    for (auto& a : next) a.store(0);
  }
};

void fail(Run& r, int rc, const std::string& what) {
  int expect = NUTDB_OK;
  if (r.rc.compare_exchange_strong(expect, rc)) {
    std::lock_guard<std::mutex> g(r.m->mu);
    r.m->err = what;
  }
}

void work(Run& r, Worker& w) {
  NutdbMCtx* m = r.m;
  if (cudaSetDevice(w.device) != cudaSuccess) {
    fail(r, NUTDB_E_CUDA, "cudaSetDevice failed");
    return;
  }
  const bool to_dev0 = (r.flags & NUTDB_MF_GATHER_DEVICE0) != 0;
  const bool want_tokens = !(r.flags & NUTDB_F_NO_TOKENS);
  cudaStream_t st = (cudaStream_t)nutdb_gpu_ctx_stream(w.ctx);
  const std::vector<uint64_t>& mine = r.per_device[w.device_index];
  for (;;) {
    if (r.rc.load() != NUTDB_OK) return;
    const uint64_t k = r.next[w.device_index].fetch_add(1);
    if (k >= mine.size()) return;
    const NutdbMShard& sh = r.shards[mine[k]];
    NutdbBatch b;
    const uint32_t pflags = (sh.flags & (NUTDB_F_DEVICE_INPUT | NUTDB_F_OFFSETS32)) | (r.flags & (NUTDB_F_NO_TOKENS | NUTDB_F_WIRE_STMT)) | NUTDB_F_NO_HOST_COPY;
    int rc = nutdb_gpu_parse_batch(w.ctx, sh.sql, sh.stmt_off, sh.n_stmt, pflags, &b);
    if (rc != NUTDB_OK) {
      fail(r, rc, std::string("shard ") + std::to_string(mine[k]) + ": " + nutdb_gpu_last_error(w.ctx));
      return;
    }
    NutdbBatchDevice dv;
    nutdb_gpu_batch_device(&b, &dv);
    if (b.n_ext && nutdb_gpu_batch_fetch_ext(&b) != NUTDB_OK) {  // (rare: nodes that did not fit the 32-bit wire word)
      fail(r, NUTDB_E_CUDA, "cannot fetch the wire nodes' side table");
      return;
    }
    // ---- gather: one copy per array from the producing device to the gather point ----
    Slot& s = to_dev0 ? w.dev0 : w.host;
    const bool wire_stmt = dv.wstmt != nullptr;  // (NUTDB_F_WIRE_STMT, and every count fitted the 8-byte record)
    const void* src[7] = {wire_stmt ? dv.wstmt : dv.stmt, dv.node, dv.err, dv.tok_type, dv.tok_start, dv.tok_end, dv.tok_kw};
    const size_t bytes[7] = {(wire_stmt ? sizeof(uint64_t) : sizeof(NutdbStmt)) * (size_t)b.n_stmt, sizeof(uint32_t) * (size_t)b.n_node,
                             sizeof(NutdbError) * (size_t)b.n_err, want_tokens ? (size_t)b.n_tok : 0,
                             want_tokens ? 4 * (size_t)b.n_tok : 0, want_tokens ? 4 * (size_t)b.n_tok : 0,
                             want_tokens ? (size_t)b.n_tok : 0};
    bool ok = true;
    for (int i = 0; i < 7 && ok; i++) {
      if (!bytes[i]) continue;
      if (slot_reserve(m, s, i, bytes[i], to_dev0) != NUTDB_OK) {
        fail(r, NUTDB_E_NOMEM, m->err);
        ok = false;
        break;
      }
      cudaError_t e;
      if (to_dev0) e = cudaMemcpyPeerAsync(s.p[i], m->devices[0], src[i], w.device, bytes[i], st);
      else e = cudaMemcpyAsync(s.p[i], src[i], bytes[i], cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) {
        fail(r, NUTDB_E_CUDA, std::string("gather copy failed: ") + cudaGetErrorString(e));
        ok = false;
      }
    }
    if (!ok) return;
    if (cudaStreamSynchronize(st) != cudaSuccess) {
      fail(r, NUTDB_E_CUDA, "gather copy failed (synchronize)");
      return;
    }
    NutdbMChunk c;
    std::memset(&c, 0, sizeof(c));
    c.shard = mine[k];
    c.first_stmt = sh.first_stmt;
    c.device = w.device;
    c.on_device = to_dev0 ? 1 : 0;
    c.batch.n_stmt = b.n_stmt;
    c.batch.n_tok = b.n_tok;
    c.batch.n_node = b.n_node;
    c.batch.n_err = b.n_err;
    if (wire_stmt) c.batch.wstmt = (const uint64_t*)(bytes[0] ? s.p[0] : nullptr);
    else c.batch.stmt = (const NutdbStmt*)(bytes[0] ? s.p[0] : nullptr);
    c.batch.pnode = (const uint32_t*)(bytes[1] ? s.p[1] : nullptr);
    c.batch.err = (const NutdbError*)(bytes[2] ? s.p[2] : nullptr);
    if (want_tokens && b.n_tok) {
      c.batch.tok_type = (const uint8_t*)s.p[3];
      c.batch.tok_start = (const uint32_t*)s.p[4];
      c.batch.tok_end = (const uint32_t*)s.p[5];
      c.batch.tok_kw = (const uint8_t*)s.p[6];
    }
    c.batch.n_ext = b.n_ext;
    c.batch.ext = b.ext;
    if (r.fn) {
      if (r.flags & NUTDB_MF_SERIAL_CALLBACKS) {
        std::lock_guard<std::mutex> g(r.cb_mu);
        r.fn(r.user, &c);
      } else {
        r.fn(r.user, &c);
      }
    }
    nutdb_gpu_batch_free(w.ctx, &b);
  }
}

}  // namespace

extern "C" {

NutdbMCtx* nutdb_gpu_mctx_create(const int* devices, int n_devices, int workers_per_device) {
  if (!devices || n_devices < 1 || n_devices > 64) return nullptr;
  if (workers_per_device < 1) workers_per_device = 3;
  if (workers_per_device > 16) workers_per_device = 16;
  NutdbMCtx* m = new (std::nothrow) NutdbMCtx();
  if (!m) return nullptr;
  m->devices.assign(devices, devices + n_devices);
  m->workers_per_device = workers_per_device;
  m->workers.resize((size_t)n_devices * workers_per_device);
  for (int d = 0; d < n_devices; d++)
    for (int k = 0; k < workers_per_device; k++) {
      Worker& w = m->workers[(size_t)d * workers_per_device + k];
      w.device_index = d;
      w.device = devices[d];
      w.ctx = nutdb_gpu_ctx_create(devices[d]);
      if (!w.ctx) {
        nutdb_gpu_mctx_destroy(m);
        return nullptr;
      }
    }
  return m;
}

void nutdb_gpu_mctx_destroy(NutdbMCtx* m) {
  if (!m) return;
  for (Worker& w : m->workers) {
    slot_free(m, w.host);
    slot_free(m, w.dev0);
    if (w.ctx) nutdb_gpu_ctx_destroy(w.ctx);
  }
  delete m;
}

const char* nutdb_gpu_mctx_last_error(const NutdbMCtx* m) { return m ? m->err.c_str() : "no dispatcher"; }
int nutdb_gpu_mctx_device_count(const NutdbMCtx* m) { return m ? (int)m->devices.size() : 0; }

int nutdb_gpu_copy_to_host(void* dst, const void* src_device, uint64_t bytes) {
  if (!bytes) return NUTDB_OK;
  if (!dst || !src_device) return NUTDB_E_ARG;
  return cudaMemcpy(dst, src_device, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? NUTDB_OK : NUTDB_E_CUDA;
}

int nutdb_gpu_mctx_parse_shards(NutdbMCtx* m, const NutdbMShard* shards, uint64_t n_shards, uint32_t flags, nutdb_chunk_fn fn,
                                void* user) {
  if (!m) return NUTDB_E_CUDA;
  if (n_shards && !shards) {
    m->err = "null argument";
    return NUTDB_E_ARG;
  }
  const size_t ndev = m->devices.size();
  Run r(ndev);
  r.m = m;
  r.shards = shards;
  r.n_shards = n_shards;
  r.flags = flags;
  r.fn = fn;
  r.user = user;
  for (uint64_t i = 0; i < n_shards; i++) {
    if (shards[i].device_index < 0 || (size_t)shards[i].device_index >= ndev) {
      m->err = "shard names a device the dispatcher does not own";
      return NUTDB_E_ARG;
    }
    r.per_device[(size_t)shards[i].device_index].push_back(i);
  }
  if ((flags & NUTDB_MF_GATHER_DEVICE0) && !m->peer_checked) {  // peer access for the NVLink gather (once)
    for (size_t d = 1; d < ndev; d++) {
      if (m->devices[d] == m->devices[0]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->devices[d], m->devices[0]);
      if (can) {
        cudaSetDevice(m->devices[d]);
        cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      }  // (without peer access cudaMemcpyPeerAsync stages through the host: slower, still correct)
    }
    m->peer_checked = true;
  }
  std::vector<std::thread> th;
  th.reserve(m->workers.size());
  for (Worker& w : m->workers)
    if (!r.per_device[(size_t)w.device_index].empty()) th.emplace_back(work, std::ref(r), std::ref(w));
  for (std::thread& t : th) t.join();
  return r.rc.load();
}

int nutdb_gpu_mctx_parse_stream(NutdbMCtx* m, const uint8_t* sql, const uint64_t* stmt_off, uint64_t n_stmt, uint64_t chunk_bytes,
                                uint32_t flags, nutdb_chunk_fn fn, void* user) {
  if (!m) return NUTDB_E_CUDA;
  if (!stmt_off || (n_stmt && !sql)) {
    m->err = "null argument";
    return NUTDB_E_ARG;
  }
  if (flags & NUTDB_F_DEVICE_INPUT) {
    m->err = "nutdb_gpu_mctx_parse_stream takes a host batch (device-resident text: nutdb_gpu_mctx_parse_shards)";
    return NUTDB_E_ARG;
  }
  // (the offsets are validated where they are used: every shard's by k_prep on its device -- walking 8 bytes per
  // statement here would cost more host time than a chunk's kernels; a descending pair either trips the range checks
  // below or makes its shard fail with NUTDB_E_ARG)
  const bool o32 = (flags & NUTDB_F_OFFSETS32) != 0;
  const uint32_t* const off32p = reinterpret_cast<const uint32_t*>(stmt_off);
  auto off = [&](uint64_t i) -> uint64_t { return o32 ? (uint64_t)off32p[i] : stmt_off[i]; };
  if (off(n_stmt) < off(0)) {
    m->err = "statement offsets must ascend";
    return NUTDB_E_ARG;
  }
  if (chunk_bytes < 4096) chunk_bytes = 4096;
  if (chunk_bytes > 0x70000000ull) chunk_bytes = 0x70000000ull;
  const uint64_t total = off(n_stmt) - off(0);
  const size_t ndev = m->devices.size();
  // cut points: statement boundaries nearest to the multiples of `chunk_bytes`; chunk k goes to device k * ndev / nchunks
  // (contiguous ranges of the batch per device, balanced by bytes)
  uint64_t nchunks = std::max<uint64_t>(1, (total + chunk_bytes - 1) / chunk_bytes);
  nchunks = std::max<uint64_t>(nchunks, std::min<uint64_t>(ndev, std::max<uint64_t>(n_stmt, 1)));
  std::vector<uint64_t> cut(nchunks + 1, n_stmt);
  cut[0] = 0;
  for (uint64_t k = 1; k < nchunks; k++) {
    const uint64_t target = off(0) + (uint64_t)((unsigned __int128)total * k / nchunks);
    const uint64_t s = o32 ? (uint64_t)(std::lower_bound(off32p, off32p + n_stmt + 1, (uint32_t)std::min<uint64_t>(target, 0xFFFFFFFFull)) - off32p)
                           : (uint64_t)(std::lower_bound(stmt_off, stmt_off + n_stmt + 1, target) - stmt_off);
    cut[k] = std::min<uint64_t>(std::max<uint64_t>(s, cut[k - 1]), n_stmt);
  }
  std::vector<NutdbMShard> sh;
  sh.reserve(nchunks);
  for (uint64_t k = 0; k < nchunks; k++) {
    if (cut[k + 1] <= cut[k]) continue;
    if (off(cut[k + 1]) < off(cut[k]) || off(cut[k + 1]) - off(cut[k]) >= 0x7FFFFFFFull) {
      m->err = "statement offsets must ascend, and no chunk may exceed 2^31 bytes";
      return NUTDB_E_ARG;
    }
    NutdbMShard s;
    s.device_index = (int)(k * ndev / nchunks);
    s.sql = sql;
    s.stmt_off = o32 ? reinterpret_cast<const uint64_t*>(off32p + cut[k]) : stmt_off + cut[k];
    s.n_stmt = cut[k + 1] - cut[k];
    s.flags = o32 ? NUTDB_F_OFFSETS32 : 0u;
    s.first_stmt = cut[k];
    sh.push_back(s);
  }
  return nutdb_gpu_mctx_parse_shards(m, sh.data(), sh.size(), flags, fn, user);
}

}  // extern "C"
