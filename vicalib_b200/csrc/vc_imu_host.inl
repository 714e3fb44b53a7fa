// Host-side orchestration of the inertial path (included by vcgpu.cu after its helper macros).

struct ImuDevHost {
  double* cost = nullptr;   // [ni] per-interval robust cost
  double* Cg = nullptr;     // [nf][kImuCgStride] per-interval global blocks
  double* ftime = nullptr;  // [nf]
  vc::imu::ImuBuf buf{};
  std::vector<vc::ChainLevel> levels;
  vc::ChainLevel* d_levels = nullptr;  // device copy for the persistent kernel
  double* pool = nullptr;
  int32_t* ipool = nullptr;
  double* Ssum = nullptr;
  int n_part = 0;
};
namespace vc { using ImuDev = ::ImuDevHost; }

constexpr int kChainC = 4;    // chunk length: every 4th node of a level is a separator
constexpr int kChainTop = 4;  // levels with <= this many nodes join the dense solve

static vc::ImuDev* imu_dev(vcgpu_handle* h) { return static_cast<vc::ImuDev*>(h->imu); }

static void imu_free(vcgpu_handle* h) {
  vc::ImuDev* d = imu_dev(h);
  if (!d) return;
  dev_free(&d->cost); dev_free(&d->Cg); dev_free(&d->ftime); dev_free(&d->pool); dev_free(&d->ipool);
  dev_free(&d->Ssum); dev_free(&d->d_levels);
  delete d;
  h->imu = nullptr;
}

static int imu_prepare(vcgpu_handle* h) {
  const DevProblem& dp = h->dp;
  if (!dp.inertial) return VCGPU_OK;
  if (dp.n_frames < 2) return fail(h, VCGPU_ERR_INVALID, "inertial terms need at least two frames");
  // the host-side record lives as long as the handle: its device buffers are keyed by their address in dev_alloc's
  // capacity table, so a re-upload of a same-sized problem allocates nothing
  if (!h->imu) h->imu = new vc::ImuDev();
  vc::ImuDev* d = imu_dev(h);
  const int nf = dp.n_frames, ni = nf - 1, G = dp.G, FD = 9;
  const int n = static_cast<int>(h->h_imu_t.size());
  // IMU samples SoA + the reference's running statistics (interpolation-buffer.h:70-85)
  std::vector<double> buf(7 * std::max(n, 1));
  double avg = 0.0;
  for (int i = 0; i < n; ++i) {
    buf[i] = h->h_imu_t[i];
    for (int k = 0; k < 3; ++k) {
      buf[static_cast<size_t>(1 + k) * n + i] = h->h_imu_w[3 * i + k];
      buf[static_cast<size_t>(4 + k) * n + i] = h->h_imu_a[3 * i + k];
    }
    const double dt = i > 0 ? h->h_imu_t[i] - h->h_imu_t[i - 1] : 0.0;
    avg = (avg * i + dt) / (i + 1);
  }
  VC_TRY(dev_alloc(h, &h->d_imu, buf.size()));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_imu, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->n_imu = n;
  d->buf.d = h->d_imu;
  d->buf.n = n;
  d->buf.start_time = n ? h->h_imu_t.front() : -1.0;
  d->buf.end_time = n ? h->h_imu_t.back() : -1.0;
  d->buf.average_dt = avg;
  VC_TRY(dev_alloc(h, &d->ftime, nf));
  CUDA_TRY(h, cudaMemcpyAsync(d->ftime, h->h_time.data(), nf * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  // weights start at 500*I (vicalibrator.h:616)
  std::vector<double> w(static_cast<size_t>(ni) * 81, 0.0);
  for (int k = 0; k < ni; ++k)
    for (int i = 0; i < 9; ++i) w[static_cast<size_t>(k) * 81 + i * 10] = 500.0;
  VC_TRY(dev_alloc(h, &h->d_wsqrt, w.size()));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_wsqrt, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  VC_TRY(dev_alloc(h, &h->d_imu_r, static_cast<size_t>(ni) * 9));
  VC_TRY(dev_alloc(h, &h->d_imu_J, static_cast<size_t>(ni) * 297));
  VC_TRY(dev_alloc(h, &d->cost, ni));
  VC_TRY(dev_alloc(h, &d->Cg, static_cast<size_t>(nf) * kImuCgStride));
  CUDA_TRY(h, cudaMemsetAsync(d->Cg, 0, static_cast<size_t>(nf) * kImuCgStride * sizeof(double), h->stream));
  // chain levels
  const size_t w_cols = 2 * FD + G + 1;
  // level sizes; a sharded rank reduces its chain to its first frame (+ the ghost) — the separators
  // of all ranks then meet in the all-reduced dense system
  std::vector<int> sizes;
  const int top = h->nranks > 1 ? 1 : kChainTop;
  for (int m = dp.n_own;; m = (m + kChainC - 1) / kChainC) {
    sizes.push_back(m + dp.ghost);
    if (m <= top) break;
  }
  size_t total = 0, itotal = 0;
  int n_part = 0;
  for (size_t l = 0; l < sizes.size(); ++l) {
    const size_t m = sizes[l];
    total += m * (2 * FD * FD + FD * G + FD);
    if (l > 0) total += m * (FD * FD + FD * G + FD);
    if (l + 1 < sizes.size()) { total += m * FD * w_cols; n_part += (static_cast<int>(m) - dp.ghost + kChainC - 1) / kChainC; }
    itotal += m;
  }
  VC_TRY(dev_alloc(h, &d->pool, total));
  VC_TRY(dev_alloc(h, &d->ipool, itotal));
  double* p = d->pool;
  int32_t* ip = d->ipool;
  d->levels.resize(sizes.size());
  for (size_t l = 0; l < sizes.size(); ++l) {
    const size_t m = sizes[l];
    vc::ChainLevel& L = d->levels[l];
    L.n = static_cast<int>(m);
    L.ghost = dp.ghost;
    L.A = p; p += m * FD * FD;
    L.U = p; p += m * FD * FD;
    L.E = p; p += m * FD * G;
    L.g = p; p += m * FD;
    L.addA = L.addE = L.addg = nullptr;
    if (l > 0) {
      L.addA = p; p += m * FD * FD;
      L.addE = p; p += m * FD * G;
      L.addg = p; p += m * FD;
    }
    L.Z = nullptr;
    if (l + 1 < sizes.size()) { L.Z = p; p += m * FD * w_cols; }
    L.orig = ip; ip += m;
  }
  d->n_part = n_part;
  const size_t NS = static_cast<size_t>(G) * G + G;
  if (static_cast<size_t>(std::max(n_part, 1)) > static_cast<size_t>(h->n_solve_blocks)) {
    VC_TRY(dev_alloc(h, &h->d_Spart, static_cast<size_t>(n_part) * NS));
  }
  VC_TRY(dev_alloc(h, &d->Ssum, NS));
  return VCGPU_OK;
}

static int wts_join(vcgpu_handle* h);
static int kernel_smem_optin(vcgpu_handle* h);
static int imu_evaluate(vcgpu_handle* h, int which, bool apply_loss, int* n_cost, const double* mask_dev = nullptr) {
  VC_TRY(wts_join(h));  // the residuals are weighted
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  const int ni = dp.n_frames - 1;
  ImuEvalArgs a;
  a.dp = dp; a.buf = d->buf; a.ctl = h->d_ctl; a.which = which;
  a.states[0] = h->d_state[0]; a.states[1] = h->d_state[1];
  a.ftime = d->ftime; a.wsqrt = h->d_wsqrt;
  a.mask = (mask_dev ? mask_dev : h->d_mask) + dp.imu_goff;
  a.r = h->d_imu_r; a.J = h->d_imu_J; a.cost = d->cost; a.ni = ni; a.apply_loss = apply_loss ? 1 : 0;
  a.mult = dp.imu_mult;
  imu_eval_kernel<<<(ni + kImuWarps - 1) / kImuWarps, 32 * kImuWarps, 0, h->stream>>>(a);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  *n_cost = ni;
  return VCGPU_OK;
}
static const double* imu_cost_part(vcgpu_handle* h) { return imu_dev(h) ? imu_dev(h)->cost : nullptr; }
static const double* imu_cg(vcgpu_handle* h) { return imu_dev(h) ? imu_dev(h)->Cg : nullptr; }

static int imu_accumulate(vcgpu_handle* h, int which) {
  vc::ImuDev* d = imu_dev(h);
  ImuAccArgs a;
  a.dp = h->dp; a.ctl = h->d_ctl; a.which = which; a.r = h->d_imu_r; a.J = h->d_imu_J;
  a.outs[0] = h->blk[0]; a.outs[1] = h->blk[1]; a.Cg = d->Cg; a.ni = h->dp.n_frames - 1;
  imu_accumulate_kernel<<<h->dp.n_frames, 128, 0, h->stream>>>(a);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}
// forward elimination of the frame chain; the summed Schur partials land in Ssum
static int imu_chain_eliminate(vcgpu_handle* h, const double* D2x) {
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  constexpr int FD = 9;
  const int G = dp.G;
  const size_t NS = static_cast<size_t>(G) * G + G;
  chain_init_kernel<FD><<<dp.n_frames, 128, 0, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, D2x, d->levels[0],
                                                           h->nranks > 1 ? h->d_sep : nullptr);
  ++h->launches;
  const size_t w_cols = 2 * FD + G + 1;
  const size_t esm = (NS + 3 * FD * FD + FD * G + FD + static_cast<size_t>(kChainC - 1) * FD * (FD + w_cols)) * sizeof(double);
  VC_TRY(kernel_smem_optin(h));
  int part = 0;
  for (size_t l = 0; l + 1 < d->levels.size(); ++l) {
    ElimArgs ea;
    ea.G = G; ea.c = kChainC; ea.ctl = h->d_ctl; ea.cur = d->levels[l]; ea.next = d->levels[l + 1];
    ea.Spart = h->d_Spart + static_cast<size_t>(part) * NS; ea.scalars = h->d_scalars;
    const int nsep = d->levels[l + 1].n - dp.ghost;  // one CTA per separator-led chunk
    chain_eliminate_kernel<FD><<<nsep, kChainThreads, esm, h->stream>>>(ea);
    ++h->launches;
    part += nsep;
  }
  sum_partials_kernel<<<static_cast<int>((NS + 31) / 32), 256, 0, h->stream>>>(h->d_Spart, part, static_cast<int>(NS), d->Ssum, h->d_ctl);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

// dense solve of [globals | top-level nodes]; globals and top-node steps land in d_delta.
// Sharded: every rank assembles its partial of the [globals | one separator per rank] system, one
// all-reduce sums them, every rank factors the same matrix.
static int all_reduce(vcgpu_handle* h, double* buf, size_t n);
static int imu_chain_dense(vcgpu_handle* h, const double* D2x) {
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  constexpr int FD = 9;
  DenseArgs da;
  da.dp = dp; da.bs[0] = h->blk[0]; da.bs[1] = h->blk[1]; da.ctl = h->d_ctl; da.scale = h->d_scale; da.D2x = D2x;
  da.Ssum = d->Ssum; da.top = d->levels.back(); da.delta = h->d_delta; da.scalars = h->d_scalars;
  da.buf = h->d_dense;
  const bool multi = h->nranks > 1;
  da.n_slots = multi ? h->nranks : da.top.n;
  for (int t = 0; t < 4; ++t) da.slot_of[t] = multi ? h->rank + t : t;  // own first frame, then the ghost
  da.add_globals = h->rank == 0 ? 1 : 0;
  const size_t N = dp.G + static_cast<size_t>(da.n_slots) * FD;
  const size_t sm = (N * N + N) * sizeof(double);
  if (!multi) {
    da.mode = 0;
    dense_solve_kernel<FD><<<1, 256, sm, h->stream>>>(da);
    ++h->launches;
  } else {
    da.mode = 1;
    dense_solve_kernel<FD><<<1, 256, 0, h->stream>>>(da);
    VC_TRY(all_reduce(h, h->d_dense, N * N + N));
    da.mode = 2;
    dense_solve_kernel<FD><<<1, 256, sm, h->stream>>>(da);
    h->launches += 2;
  }
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

// back-substitution through the levels (top-down), then x (+) delta
static int imu_chain_backsub(vcgpu_handle* h, const double* D2x, bool explicit_update) {
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  constexpr int FD = 9;
  for (int l = static_cast<int>(d->levels.size()) - 2; l >= 0; --l) {
    BacksubArgs ba;
    ba.ctl = h->d_ctl; ba.G = dp.G; ba.c = kChainC; ba.nfp_off = 0; ba.cur = d->levels[l]; ba.delta = h->d_delta;
    ba.nfp = static_cast<int64_t>(dp.n_frames) * FD;
    chain_backsub_kernel<FD><<<(d->levels[l].n + 3) / 4, 128, 0, h->stream>>>(ba);
    ++h->launches;
  }
  if (explicit_update) {  // otherwise the fused evaluate kernel applies x (+) delta itself
    UpdateArgs ua;
    ua.dp = dp; ua.b[0] = h->blk[0]; ua.b[1] = h->blk[1]; ua.ctl = h->d_ctl; ua.scale = h->d_scale; ua.D2x = D2x;
    ua.X = nullptr; ua.delta = h->d_delta; ua.state[0] = h->d_state[0]; ua.state[1] = h->d_state[1];
    ua.step_part = h->d_red; ua.sepdiag = h->nranks > 1 ? h->d_sep : nullptr;
    const int nb = (dp.n_frames + kUpdateWarps - 1) / kUpdateWarps;
    backsub_update_kernel<9><<<nb, 32 * kUpdateWarps, 0, h->stream>>>(ua);
    ++h->launches;
    h->n_step_part = nb + 1;
  }
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

// UpdateImuWeights (vicalibrator.h:723-799): active only when inertial && !rotation_only (:725)
// the side-stream weight update (if any) must be complete before the main stream goes on
static int wts_join(vcgpu_handle* h) {
  if (h->wts_pending) {
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_wts, 0));
    h->wts_pending = false;
  }
  return VCGPU_OK;
}
static int imu_update_weights(vcgpu_handle* h, bool side_stream = false, bool deferred = false) {
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  if (!dp.inertial || dp.rotation_only) return VCGPU_OK;
  VC_TRY(wts_join(h));
  side_stream = side_stream && !h->profiling && !h->flush_l2 && h->stream2 != nullptr;  // timed per iteration: stay serial
  cudaStream_t st_launch = h->stream;
  if (side_stream) {  // everything enqueued so far (the decision included) happens before the update
    CUDA_TRY(h, cudaEventRecord(h->ev_dec, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream2, h->ev_dec, 0));
    st_launch = h->stream2;
  }
  StageScope st(h, VCGPU_STAGE_IMU_WEIGHTS);
  vc::wts::WeightArgs a;
  a.dp = dp; a.buf = d->buf; a.ctl = h->d_ctl; a.states[0] = h->d_state[0]; a.states[1] = h->d_state[1];
  a.ftime = d->ftime; a.wsqrt = h->d_wsqrt;
  a.ni = dp.n_frames - 1; a.sigma_g = h->sigma_g; a.sigma_a = h->sigma_a; a.deferred = deferred ? 1 : 0;
  vc::wts::imu_weights_kernel<<<(a.ni + vc::wts::kWtTeams - 1) / vc::wts::kWtTeams, 32 * vc::wts::kWtWarps, 0, st_launch>>>(a);
  ++h->launches;
  CUDA_TRY(h, cudaGetLastError());
  if (side_stream) {
    CUDA_TRY(h, cudaEventRecord(h->ev_wts, h->stream2));
    h->wts_pending = true;
  }
  return VCGPU_OK;
}

static int ctl_reset(vcgpu_handle* h, int fixed, int max_iters);

static int imu_eval_hook(vcgpu_handle* h, double* r, double* J) {
  const DevProblem& dp = h->dp;
  if (!dp.inertial) return fail(h, VCGPU_ERR_INVALID, "eval_imu: inertial flag is off");
  const int ni = dp.n_frames - 1;
  double* ones = nullptr;
  VC_TRY(dev_alloc(h, &ones, dp.G));
  std::vector<double> hones(dp.G, 1.0);
  CUDA_TRY(h, cudaMemcpy(ones, hones.data(), dp.G * sizeof(double), cudaMemcpyHostToDevice));
  VC_TRY(ctl_reset(h, 0, 0));
  int nc = 0;
  VC_TRY(imu_evaluate(h, 0, false, &nc, ones));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(ones);
  h->blocks_valid = false;
  CUDA_TRY(h, cudaMemcpy(r, h->d_imu_r, static_cast<size_t>(ni) * 9 * sizeof(double), cudaMemcpyDeviceToHost));
  if (J) CUDA_TRY(h, cudaMemcpy(J, h->d_imu_J, static_cast<size_t>(ni) * 297 * sizeof(double), cudaMemcpyDeviceToHost));
  return VCGPU_OK;
}
