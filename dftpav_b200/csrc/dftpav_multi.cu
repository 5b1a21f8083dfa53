// dftpav_multi.cu — multi-GPU entry points of the C-ABI (include/dftpav_b200.h): one host thread + one handle per device,
// contiguous shards, and the path's single collective - an NCCL all-reduce (MIN) of the 1-int convergence flag.
//
// the batch shards embarrassingly (no term couples two trajectories; obstacles are read-only inputs),
// so there is no data-path collective.  A C++ caller (the plan manager with many candidate problems) gets from
// dftpav_solve_multi what bench.py gets from torchrun + torch.distributed: every GPU solves a contiguous slice through the
// same dftpav_solve the single-GPU path uses (results are bit-identical to a single-GPU solve of the whole batch: each
// problem's arithmetic is confined to one warp), the results land in the caller's arrays, and the devices agree on one flag.
// NCCL is loaded at run time (dlopen libnccl.so.2); without it (or with one device) the flag is reduced on the host and
// dftpav_multi_uses_nccl() says so.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dftpav_b200.h"

extern "C" int dftpav_internal_set_min_strides(dftpav_handle* h, int n_stride, int row_stride);

namespace {

// the handful of NCCL declarations used here (nccl.h: ncclResult_t, ncclComm_t, ncclInt32 = 2, ncclMin = 3)
typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    // a copy the process has already loaded (e.g. the one bundled with torch) first; otherwise a private one: RTLD_LOCAL keeps
    // its symbols away from whatever loads another NCCL later
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) return false;
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
  }
};
constexpr int kNcclInt32 = 2, kNcclMin = 3;

}  // namespace

struct dftpav_multi {
  dftpav_params prm;
  std::vector<int> devs;
  std::vector<dftpav_handle*> handles;
  std::vector<cudaStream_t> streams;
  std::vector<int*> d_flag;
  std::vector<nccl_comm_t> comms;
  NcclApi nccl;
  bool use_nccl = false;
  std::string err;
  std::vector<dftpav_stats> stats;
};

namespace {

// view of problems [p0, p1) of a batch: pointer offsets into the caller's arrays plus two small rebased prefix arrays
struct Shard {
  dftpav_batch b;
  std::vector<int32_t> seg_ptr, obs_ptr;
  int p0 = 0, p1 = 0, g0 = 0;
};

void make_shards(const dftpav_params& prm, const dftpav_batch& b, int n, std::vector<Shard>& out) {
  const int B = b.B;
  out.assign(n, Shard());
  const int K = prm.traj_resolution, Kd = prm.des_traj_resolution;
  std::vector<int> cut(n + 1);
  for (int d = 0; d <= n; d++) cut[d] = (int)((long long)B * d / n);
  long long inner = 0, samples = 0, planes = 0;
  int d = 0;
  for (int p = 0; p <= B; p++) {
    while (d < n && cut[d] == p) {
      Shard& s = out[d];
      s.p0 = cut[d], s.p1 = cut[d + 1];
      s.g0 = b.seg_ptr[s.p0];
      const int g0 = s.g0;
      s.b = b;
      s.b.B = s.p1 - s.p0;
      s.seg_ptr.resize(s.b.B + 1);
      for (int q = 0; q <= s.b.B; q++) s.seg_ptr[q] = b.seg_ptr[s.p0 + q] - g0;
      s.b.seg_ptr = s.seg_ptr.data();
      s.b.piece_num = b.piece_num + g0;
      s.b.singul = b.singul + g0;
      s.b.ini_state = b.ini_state + 6 * (size_t)g0;
      s.b.fin_state = b.fin_state + 6 * (size_t)g0;
      s.b.init_T = b.init_T + g0;
      s.b.inner_pts = b.inner_pts + inner;
      s.b.plane_cnt = b.plane_cnt ? b.plane_cnt + samples : nullptr;
      s.b.planes = b.planes + 4 * planes;
      s.b.t_now = b.t_now ? b.t_now + s.p0 : nullptr;
      if (b.obs_ptr) {  // the obstacle pool is shared by all shards; the per-problem index lists are sliced
        s.obs_ptr.resize(s.b.B + 1);
        const int o0 = b.obs_ptr[s.p0];
        for (int q = 0; q <= s.b.B; q++) s.obs_ptr[q] = b.obs_ptr[s.p0 + q] - o0;
        s.b.obs_ptr = s.obs_ptr.data();
        s.b.obs_idx = b.obs_idx ? b.obs_idx + o0 : nullptr;
      }
      d++;
    }
    if (p == B) break;
    for (int g = b.seg_ptr[p]; g < b.seg_ptr[p + 1]; g++) {
      const int N = b.piece_num[g];
      const long long M = (long long)(N - 2) * (K + 1) + 2 * (Kd + 1);
      inner += 2LL * (N - 1);
      if (b.plane_cnt)
        for (long long q = 0; q < M; q++) planes += b.plane_cnt[samples + q];
      else
        planes += M * b.uniform_H;
      samples += M;
    }
  }
}

}  // namespace

extern "C" {

int dftpav_create_multi(const dftpav_params* prm, const int32_t* device_ids, int32_t n_dev, dftpav_multi** out) {
  if (!prm || !out || n_dev < 1 || !device_ids) return DFTPAV_ERR_INVALID_ARG;
  *out = nullptr;
  dftpav_multi* m = new dftpav_multi();
  m->prm = *prm;
  m->devs.assign(device_ids, device_ids + n_dev);
  m->stats.resize(n_dev);
  for (int d = 0; d < n_dev; d++) {
    dftpav_handle* h = nullptr;
    const int rc = dftpav_create(prm, device_ids[d], &h);
    if (rc != DFTPAV_OK) {
      dftpav_destroy_multi(m);
      return rc;
    }
    m->handles.push_back(h);
    cudaSetDevice(device_ids[d]);
    cudaStream_t st = nullptr;
    int* f = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&f, sizeof(int)) != cudaSuccess) {
      dftpav_destroy_multi(m);
      return DFTPAV_ERR_CUDA;
    }
    m->streams.push_back(st);
    m->d_flag.push_back(f);
  }
  // one communicator per device, all in this process (NVLink / NVSwitch between the GPUs of the box)
  if (n_dev > 1 && m->nccl.load()) {
    m->comms.assign(n_dev, nullptr);
    if (m->nccl.CommInitAll(m->comms.data(), n_dev, m->devs.data()) == 0) {
      m->use_nccl = true;
    } else {
      m->comms.clear();
    }
  }
  *out = m;
  return DFTPAV_OK;
}

int dftpav_destroy_multi(dftpav_multi* m) {
  if (!m) return DFTPAV_ERR_INVALID_ARG;
  if (m->use_nccl)
    for (nccl_comm_t c : m->comms)
      if (c) m->nccl.CommDestroy(c);
  for (size_t d = 0; d < m->handles.size(); d++) {
    cudaSetDevice(m->devs[d]);
    if (d < m->d_flag.size() && m->d_flag[d]) cudaFree(m->d_flag[d]);
    if (d < m->streams.size() && m->streams[d]) cudaStreamDestroy(m->streams[d]);
    dftpav_destroy(m->handles[d]);
  }
  delete m;
  return DFTPAV_OK;
}

int dftpav_multi_uses_nccl(const dftpav_multi* m) { return m && m->use_nccl ? 1 : 0; }
int dftpav_multi_devices(const dftpav_multi* m) { return m ? (int)m->handles.size() : 0; }
const char* dftpav_multi_last_error(const dftpav_multi* m) { return m ? m->err.c_str() : "null handle"; }

int dftpav_multi_stats(const dftpav_multi* m, int32_t dev_index, dftpav_stats* stats) {
  if (!m || !stats || dev_index < 0 || dev_index >= (int)m->stats.size()) return DFTPAV_ERR_INVALID_ARG;
  *stats = m->stats[dev_index];
  return DFTPAV_OK;
}

int dftpav_solve_multi(dftpav_multi* m, const dftpav_batch* batch, const dftpav_result* out, int32_t* all_accepted) {
  if (!m || !batch || !out) return DFTPAV_ERR_INVALID_ARG;
  dftpav_dims dims;
  int rc = dftpav_batch_dims(&m->prm, batch, &dims);
  if (rc != DFTPAV_OK) {
    m->err = "invalid batch";
    return rc;
  }
  const int n = (int)m->handles.size();
  std::vector<Shard> shards;
  make_shards(m->prm, *batch, n, shards);
  std::vector<int> rcs(n, DFTPAV_OK), local_ok(n, 1);
  // ---- one host thread per device: dftpav_solve on its shard, results straight into the caller's arrays (the shard's
  //      strides are forced to the batch's so that rows line up)
  auto work = [&](int d) {
    Shard& s = shards[d];
    dftpav_handle* h = m->handles[d];
    if (s.b.B == 0) return;
    dftpav_internal_set_min_strides(h, dims.n_stride, dims.row_stride);
    dftpav_result r;
    r.x = out->x ? out->x + (size_t)s.p0 * dims.n_stride : nullptr;
    r.cost = out->cost ? out->cost + s.p0 : nullptr;
    r.status = out->status ? out->status + s.p0 : nullptr;
    r.success = out->success ? out->success + s.p0 : nullptr;
    r.iters = out->iters ? out->iters + s.p0 : nullptr;
    r.evals = out->evals ? out->evals + s.p0 : nullptr;
    r.coeffs = out->coeffs ? out->coeffs + (size_t)s.p0 * dims.row_stride * 2 : nullptr;
    r.piece_dT = out->piece_dT ? out->piece_dT + s.g0 : nullptr;
    rcs[d] = dftpav_solve(h, &s.b, &r);
    if (rcs[d] == DFTPAV_OK) {
      int32_t ok = 0;
      rcs[d] = dftpav_success_count(h, &ok);
      local_ok[d] = (ok == s.b.B) ? 1 : 0;
      dftpav_last_stats(h, &m->stats[d]);
    }
  };
  std::vector<std::thread> threads;
  for (int d = 1; d < n; d++) threads.emplace_back(work, d);
  work(0);
  for (std::thread& t : threads) t.join();
  for (int d = 0; d < n; d++)
    if (rcs[d] != DFTPAV_OK) {
      m->err = std::string("device ") + std::to_string(m->devs[d]) + ": " + dftpav_last_error(m->handles[d]);
      return rcs[d];
    }
  // ---- the single collective: MIN over the devices of "every problem of my shard finished and was accepted"
  int flag = 1;
  if (m->use_nccl) {
    for (int d = 0; d < n; d++) {
      cudaSetDevice(m->devs[d]);
      cudaMemcpyAsync(m->d_flag[d], &local_ok[d], sizeof(int), cudaMemcpyHostToDevice, m->streams[d]);
    }
    m->nccl.GroupStart();
    int nrc = 0;
    for (int d = 0; d < n; d++)
      nrc |= m->nccl.AllReduce(m->d_flag[d], m->d_flag[d], 1, kNcclInt32, kNcclMin, m->comms[d], m->streams[d]);
    nrc |= m->nccl.GroupEnd();
    if (nrc != 0) {
      m->err = "ncclAllReduce failed";
      return DFTPAV_ERR_CUDA;
    }
    for (int d = 0; d < n; d++) {
      cudaSetDevice(m->devs[d]);
      int v = 0;
      cudaMemcpyAsync(&v, m->d_flag[d], sizeof(int), cudaMemcpyDeviceToHost, m->streams[d]);
      if (cudaStreamSynchronize(m->streams[d]) != cudaSuccess) {
        m->err = "flag read-back failed";
        return DFTPAV_ERR_CUDA;
      }
      if (d == 0) flag = v;
      else if (v != flag) {
        m->err = "devices disagree on the reduced flag";
        return DFTPAV_ERR_CUDA;
      }
    }
  } else {
    for (int d = 0; d < n; d++) flag = std::min(flag, local_ok[d]);
  }
  if (all_accepted) *all_accepted = flag;
  return DFTPAV_OK;
}

}  // extern "C"
