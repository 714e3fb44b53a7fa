// The iteration engine (included by vcgpu.cu): kernel launch helpers, the evaluation pass, the
// damped arrow solve and the trust-region loop.  All trust-region state lives in a device-resident
// Ctl block.  Vision solves (one GPU or frame shards) run in the persistent kernel (mega_launch,
// vc_mega.cuh): one launch per solve.  Everything else — inertial solves, DOGLEG, the inspection hooks,
// profiling — is a fixed sequence of launches the host enqueues without waiting:
//
//   frame_solve -> sum_partials -> global_solve -> backsub_update        (vision)
//   chain_init -> chain_eliminate x levels -> sum_partials -> dense_solve -> chain_backsub x levels
//              -> backsub_update                                         (inertial)
//   fused_build [-> imu_eval -> imu_accumulate] -> reduce_finalize (+ decide_step) [-> imu_weights, side stream]
//
// Multi-GPU on this path: an NCCL all-reduce of the reduced system after the partial sums and one of the
// global blocks / scalars before the decision are the only cross-rank points (frames are sharded; every
// rank then solves the same small dense system).

// Measured on B200 (config 2): summing the Schur partials inside the single-CTA dense solve is
// latency-bound (+18 us); it stays a separate, fully parallel launch.
constexpr bool kSumInSolve = false;

// ------------------------------------------------------------------ multi-GPU all-reduce
static int all_reduce(vcgpu_handle* h, double* buf, size_t n) {
  if (h->nranks <= 1) return VCGPU_OK;
  const ncclResult_t rc = ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, static_cast<ncclComm_t>(h->comm), h->stream);
  if (rc != ncclSuccess) return fail(h, VCGPU_ERR_COMM, std::string("ncclAllReduce: ") + ncclGetErrorString(rc));
  ++h->collectives;
  return VCGPU_OK;
}

// ------------------------------------------------------------------ control block
static int ctl_upload(vcgpu_handle* h) {
  CUDA_TRY(h, cudaMemcpyAsync(h->d_ctl, h->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
  return VCGPU_OK;
}
static int ctl_download(vcgpu_handle* h) {
  CUDA_TRY(h, cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  stage_collect(h);
  h->cur = h->h_ctl->cur;
  return VCGPU_OK;
}
// fresh control block for an inspection hook or a new solve
static int ctl_reset(vcgpu_handle* h, int fixed, int max_iters) {
  Ctl* c = h->h_ctl;
  std::memset(c, 0, sizeof *c);
  c->cur = h->cur;
  c->fixed = fixed;
  c->max_iters = max_iters;
  c->radius = h->opts.init_radius;
  c->decrease_factor = 2.0;
  c->dl_mu = 1e-8;  // DoglegStrategy: mu_ = min_mu_
  c->dl_ok = 1;
  c->function_tol = h->opts.function_tol;
  c->gradient_tol = h->opts.gradient_tol;
  c->param_tol = h->opts.param_tol;
  return ctl_upload(h);
}

// ------------------------------------------------------------------ reprojection pass (two-pass path + hooks)
template <bool JAC>
static void launch_eval_cam(vcgpu_handle* h, const EvalArgs& a, int model, int nblk) {
  switch (model) {
    case kLinear: eval_reproj_kernel<kLinear, JAC><<<nblk, 256, 0, h->stream>>>(a); break;
    case kFov: eval_reproj_kernel<kFov, JAC><<<nblk, 256, 0, h->stream>>>(a); break;
    case kPoly2: eval_reproj_kernel<kPoly2, JAC><<<nblk, 256, 0, h->stream>>>(a); break;
    case kPoly3: eval_reproj_kernel<kPoly3, JAC><<<nblk, 256, 0, h->stream>>>(a); break;
    default: eval_reproj_kernel<kKb4, JAC><<<nblk, 256, 0, h->stream>>>(a); break;
  }
  ++h->launches;
}

static int eval_reproj(vcgpu_handle* h, int which, bool jac, bool apply_loss, const double* mask_dev) {
  VC_TRY(ensure_rJ(h));
  const DevProblem& dp = h->dp;
  const int64_t n = h->n_obs;
  int part = 0;
  for (int c = 0; c < dp.n_cams; ++c) {
    const CamInfo& ci = dp.cams[c];
    if (ci.n_obs == 0) continue;
    EvalArgs a;
    a.state[0] = h->d_state[0]; a.state[1] = h->d_state[1]; a.ctl = h->d_ctl; a.which = which;
    a.cam_off = dp.off_cam + kCamStateStride * c;
    a.frame = h->d_obs_frame + ci.obs_start;
    a.pw = h->d_pw + 3 * static_cast<int64_t>(ci.obs_start);
    a.pc = h->d_pc + 2 * static_cast<int64_t>(ci.obs_start);
    a.mask = mask_dev + ci.goff;
    a.r0 = h->d_r + ci.obs_start;
    a.r1 = h->d_r + n + ci.obs_start;
    a.J = h->d_J + ci.joff;
    a.cost_part = h->d_cost_part + part;
    a.n = ci.n_obs;
    a.apply_loss = apply_loss ? 1 : 0;
    a.mult = dp.visual_mult;
    const int nblk = (ci.n_obs + 255) / 256;
    part += nblk;
    if (jac) launch_eval_cam<true>(h, a, ci.model, nblk);
    else launch_eval_cam<false>(h, a, ci.model, nblk);
  }
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

// Opt-in to more than 48 KB of dynamic shared memory for every kernel of the multi-launch engine that is launched with a
// G-dependent amount (8 cameras + IMU: NS = G^2 + G = 20 592 doubles = 161 KB).  The attribute is per device and function:
// set once per handle (a process may hold handles on several devices).
static int kernel_smem_optin(vcgpu_handle* h) {
  if (h->smem_optin_done) return VCGPU_OK;
  const int big = 200 * 1024;
  CUDA_TRY(h, cudaFuncSetAttribute(fused_build_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kFusedSmemDoubles * sizeof(double))));
  CUDA_TRY(h, cudaFuncSetAttribute(fused_build_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kFusedSmemDoubles * sizeof(double))));
  CUDA_TRY(h, cudaFuncSetAttribute(frame_solve_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_TRY(h, cudaFuncSetAttribute(reduce_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_TRY(h, cudaFuncSetAttribute(global_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_TRY(h, cudaFuncSetAttribute(chain_eliminate_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CUDA_TRY(h, cudaFuncSetAttribute(dense_solve_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  h->smem_optin_done = true;
  return VCGPU_OK;
}

// ------------------------------------------------------------------ persistent inertial kernels
// chain_solve_kernel + eval_mega_kernel (vc_imu_mega.cuh, vc_imu_eval_mega.cuh): two cooperative launches per iteration.
// Frame-sharded runs use them too when the ranks' totals buffers are mapped (vcgpu_comm_init): the two reductions of
// an iteration then go through the in-kernel exchange (vc_xchg.cuh) instead of NCCL launches between kernels.
static bool imu_mega_applies(const vcgpu_handle* h) {
  return h->imu_mega_ok && h->dp.inertial && (h->nranks == 1 || h->xchg_ready) && !h->materialize && !h->multi_launch;
}
static void imu_mega_xchg(const vcgpu_handle* h, Xchg* x, size_t off, int stride, unsigned tag) {
  x->rank = h->rank; x->nranks = h->nranks; x->off = off; x->stride = stride; x->tag = tag; x->ctl = h->d_ctl;
  for (int r = 0; r < kMaxRanks; ++r) x->buf[r] = reinterpret_cast<unsigned long long*>(h->xchg_peer[r]);
}
static int imu_mega_dense_stride(const DevProblem& dp) { return dp.G * dp.G + dp.G + 2 * chain_top_block(dp.G); }
static int imu_mega_eval_stride(const DevProblem& dp) { return dp.G * dp.G + dp.G + 8 + 4 * 9; }
static int imu_mega_prepare(vcgpu_handle* h) {
  const DevProblem& dp = h->dp;
  h->imu_mega_ok = false;
  if (!dp.inertial || (h->nranks > 1 && !h->xchg_ready)) return VCGPU_OK;
  if (h->nranks > 1) {  // the exchange regions must hold [2 parities][ranks][stride] entries of two words
    const size_t nd = 4 * static_cast<size_t>(h->nranks) * imu_mega_dense_stride(dp), ne = 4 * static_cast<size_t>(h->nranks) * imu_mega_eval_stride(dp);
    if (kXchgImuDenseOff + nd > kXchgImuEvalOff || kXchgImuEvalOff + ne > kXchgWords) return VCGPU_OK;
  }
  if (h->dev_sms == 0) {
    int coop = 0, smem_optin = 0, sms = 0;
    CUDA_TRY(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    CUDA_TRY(h, cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    h->dev_sms = sms;
    h->dev_smem_optin = coop ? smem_optin : 0;
  }
  const size_t sm_solve = chain_solve_smem_doubles(dp.G, h->nranks) * sizeof(double);
  const size_t sm_eval = eval_mega_smem_doubles(dp.G) * sizeof(double);
  vc::ImuDev* d = imu_dev(h);
  if (h->dev_smem_optin == 0 || sm_solve > static_cast<size_t>(h->dev_smem_optin) || sm_eval > static_cast<size_t>(h->dev_smem_optin) ||
      d->levels.size() > static_cast<size_t>(kMaxChainLevels) || 3 * 9 + dp.G + 1 > kCsGroup)
    return VCGPU_OK;  // does not fit: the multi-launch engine runs the solve
  CUDA_TRY(h, cudaFuncSetAttribute(chain_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_solve)));
  CUDA_TRY(h, cudaFuncSetAttribute(eval_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_eval)));
  int per_sm = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_solve_kernel, kCsThreads, sm_solve));
  if (per_sm < 1) return VCGPU_OK;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eval_mega_kernel, kEvThreads, sm_eval));
  if (per_sm < 1) return VCGPU_OK;
  const int grid = h->dev_sms;
  const size_t NS = static_cast<size_t>(dp.G) * dp.G + dp.G;
  VC_TRY(dev_alloc(h, &h->d_Spart, std::max<size_t>(static_cast<size_t>(grid) * kCsGroups, std::max(h->n_solve_blocks, d->n_part)) * NS));
  VC_TRY(dev_alloc(h, &h->d_Cpart, std::max<size_t>(grid, kReduceBlocks) * NS));
  VC_TRY(dev_alloc(h, &h->d_red_part, 8 * std::max<size_t>(grid, kReduceBlocks)));
  VC_TRY(dev_alloc(h, &h->d_red, 4 * (std::max<size_t>(dp.n_frames, grid) + 2)));
  VC_TRY(dev_alloc(h, &h->d_prof2, kImuProfSlots));
  CUDA_TRY(h, cudaMemsetAsync(h->d_prof2, 0, kImuProfSlots * sizeof(unsigned long long), h->stream));
  if (h->nranks > 1) VC_TRY(dev_alloc(h, &h->d_dsys, NS + static_cast<size_t>(h->nranks) * chain_top_block(dp.G)));
  if (!h->d_csync) {
    CUDA_TRY(h, cudaMalloc(&h->d_csync, 8 * sizeof(unsigned long long)));
    CUDA_TRY(h, cudaMemsetAsync(h->d_csync, 0, 8 * sizeof(unsigned long long), h->stream));
  }
  VC_TRY(dev_alloc(h, &d->d_levels, d->levels.size()));
  CUDA_TRY(h, cudaMemcpyAsync(d->d_levels, d->levels.data(), d->levels.size() * sizeof(vc::ChainLevel), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->imu_mega_grid = grid;
  h->imu_mega_ok = true;
  return VCGPU_OK;
}
static int imu_mega_solve(vcgpu_handle* h, const double* D2x, bool do_update, bool deferred_weights = false) {
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  ChainSolveArgs ca;
  ca.dp = dp; ca.b[0] = h->blk[0]; ca.b[1] = h->blk[1]; ca.ctl = h->d_ctl; ca.scale = h->d_scale; ca.D2x = D2x;
  ca.n_levels = static_cast<int>(d->levels.size());
  ca.lev = d->d_levels;
  ca.Spart = h->d_Spart; ca.Ssum = d->Ssum; ca.delta = h->d_delta; ca.scalars = h->d_scalars;
  ca.state[0] = h->d_state[0]; ca.state[1] = h->d_state[1]; ca.step_part = h->d_red; ca.do_update = do_update ? 1 : 0;
  ca.prof = (h->phase_clocks || h->profiling) ? h->d_prof2 : nullptr;
  ca.narrow_ok = std::getenv("VCGPU_NO_NARROW") ? 0 : 1;
  imu_mega_xchg(h, &ca.x, kXchgImuDenseOff, imu_mega_dense_stride(dp), h->nranks > 1 ? ++h->xchg_tag_dense : 0u);
  ca.sepdiag = h->nranks > 1 ? h->d_sep : nullptr;
  ca.dsys = h->d_dsys;
  // CTAs that stay with the solve while the others work the weights queue: as many as leave one 16-interval task per
  // leaving CTA (the update then takes one round), between 16 and 32 (measured on the target workload: 10..20 within
  // 2 %, 24 — one task too many for one round — 14 % slower)
  static const int n_solver_env = std::getenv("VCGPU_N_SOLVER") ? std::atoi(std::getenv("VCGPU_N_SOLVER")) : 0;
  const int wtasks = (dp.n_frames - 1 + 15) / 16;
  ca.wts_on = deferred_weights ? 1 : 0;
  ca.n_solver = n_solver_env > 0 ? n_solver_env : std::min(32, std::max(16, h->imu_mega_grid - wtasks - 4));
  ca.buf = d->buf; ca.ftime = d->ftime; ca.wsqrt = h->d_wsqrt; ca.sigma_g = h->sigma_g; ca.sigma_a = h->sigma_a;
  const int par = static_cast<int>(h->cs_launches++ & 1u);
  ca.sync = h->d_csync + 4 * par; ca.sync_next = h->d_csync + 4 * (1 - par);
  void* args[] = {&ca};
  CUDA_TRY(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(chain_solve_kernel), dim3(h->imu_mega_grid), dim3(kCsThreads), args,
                                          chain_solve_smem_doubles(dp.G, h->nranks) * sizeof(double), h->stream));
  ++h->launches;
  if (do_update) h->n_step_part = h->imu_mega_grid + 1;
  return VCGPU_OK;
}
static int wts_join(vcgpu_handle* h);
static int imu_mega_eval(vcgpu_handle* h, int which, bool with_step, int decide_mode, bool weights) {
  VC_TRY(wts_join(h));
  vc::ImuDev* d = imu_dev(h);
  const DevProblem& dp = h->dp;
  const bool visual = h->flags.visual && h->n_obs > 0;
  EvalMegaArgs ea;
  ea.dp = dp;
  if (!visual) ea.dp.n_cams = 0;
  ea.ctl = h->d_ctl; ea.which = which; ea.decide_mode = decide_mode; ea.do_weights = weights ? 1 : 0;
  ea.state[0] = h->d_state[0]; ea.state[1] = h->d_state[1]; ea.blk[0] = h->blk[0]; ea.blk[1] = h->blk[1];
  ea.grp_start = h->d_grp_start; ea.grp_count = h->d_grp_count; ea.group_of = h->d_group_of;
  ea.pw = h->d_pw; ea.pc = h->d_pc; ea.mask = h->d_mask; ea.Cg = h->d_Cg; ea.cost_part = h->d_cost_part;
  ea.buf = d->buf; ea.ftime = d->ftime; ea.wsqrt = h->d_wsqrt; ea.imu_r = h->d_imu_r; ea.imu_J = h->d_imu_J;
  ea.imu_cost = d->cost; ea.imuCg = d->Cg; ea.sigma_g = h->sigma_g; ea.sigma_a = h->sigma_a;
  ea.Cpart = h->d_Cpart; ea.red_part = h->d_red_part;
  // the last step_part slot is the globals' share: counted once (rank 0) in a sharded run
  ea.step_part = with_step ? h->d_red : nullptr; ea.n_step_part = h->n_step_part - (h->rank == 0 ? 0 : 1);
  ea.scalars = h->d_scalars; ea.counter = h->d_counter + 2;
  imu_mega_xchg(h, &ea.x, kXchgImuEvalOff, imu_mega_eval_stride(dp), h->nranks > 1 ? ++h->xchg_tag_eval : 0u);
  ea.sep_out = h->d_sep;
  ea.prof = (h->phase_clocks || h->profiling) ? h->d_prof2 + 32 : nullptr;
  void* args[] = {&ea};
  CUDA_TRY(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(eval_mega_kernel), dim3(h->imu_mega_grid), dim3(kEvThreads), args,
                                          eval_mega_smem_doubles(dp.G) * sizeof(double), h->stream));
  ++h->launches;
  return VCGPU_OK;
}
// fold the persistent inertial kernels' phase clocks into the stage times (call after a stream synchronise)
static int imu_mega_collect_clocks(vcgpu_handle* h, int iters) {
  if (!(h->phase_clocks || h->profiling) || !h->d_prof2 || !h->imu_mega_ok) return VCGPU_OK;
  unsigned long long ns[kImuProfSlots];
  CUDA_TRY(h, cudaMemcpy(ns, h->d_prof2, sizeof ns, cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemset(h->d_prof2, 0, sizeof ns));
  for (int k = 0; k < kImuProfSlots; ++k) h->phase_ns[k] += ns[k];
  bool seen[VCGPU_STAGE_COUNT] = {};
  auto add = [&](int stage, unsigned long long v) { h->st_ms[stage] += v * 1e-6; seen[stage] = true; };
  for (int k = 0; k < kCsProfCount; ++k)
    add(k < kCsProfReduce ? VCGPU_STAGE_FRAME_SOLVE : k <= kCsProfDense ? VCGPU_STAGE_GLOBAL_SOLVE : k == kCsProfWeights ? VCGPU_STAGE_IMU_WEIGHTS
                                                                                                     : VCGPU_STAGE_BACKSUB, ns[k]);
  const int map_e[kEvProfCount] = {VCGPU_STAGE_EVAL_TASKS, VCGPU_STAGE_IMU_ACCUM, VCGPU_STAGE_REDUCE, VCGPU_STAGE_FINALIZE,
                                   VCGPU_STAGE_IMU_WEIGHTS};
  for (int k = 0; k < kEvProfCount; ++k) add(map_e[k], ns[32 + k]);
  for (int s = 0; s < VCGPU_STAGE_COUNT; ++s) if (seen[s]) h->st_n[s] += iters;
  return VCGPU_OK;
}

// ------------------------------------------------------------------ evaluation pass
// which = 0: the accepted point (buffers[cur]); which = 1: the trial point (buffers[1-cur]).
// Residuals, Jacobians and block normal equations land in blk[buffer]; cost / gradient norms (and
// the step reductions when with_step) land in d_scalars for decide_step.
static int evaluate_into(vcgpu_handle* h, int which, bool with_step, int decide_mode, const double* D2x = nullptr) {
  const DevProblem& dp = h->dp;
  if (imu_mega_applies(h)) return imu_mega_eval(h, which, with_step, decide_mode, false);
  const bool visual = h->flags.visual && h->n_obs > 0;
  const bool fused = !h->materialize;
  if (visual && !fused) {
    StageScope st(h, VCGPU_STAGE_EVAL_REPROJ);
    VC_TRY(eval_reproj(h, which, true, true, h->d_mask));
  }
  int n_imu_cost = 0;
  DevProblem vdp = dp;
  if (!visual) vdp.n_cams = 0;
  int n_vis_cost = visual ? h->n_cost_part : 0;
  if (fused) {
    StageScope st(h, VCGPU_STAGE_BUILD);
    FusedArgs fa;
    fa.dp = vdp; fa.ctl = h->d_ctl; fa.which = which;
    fa.state[0] = h->d_state[0]; fa.state[1] = h->d_state[1];
    fa.grp_start = h->d_grp_start; fa.grp_count = h->d_grp_count; fa.group_of = h->d_group_of;
    fa.pw = h->d_pw; fa.pc = h->d_pc; fa.n_obs = h->n_obs; fa.mask = h->d_mask;
    fa.out[0] = h->blk[0]; fa.out[1] = h->blk[1]; fa.Cg = h->d_Cg; fa.cost_part = h->d_cost_part;
    const size_t fsm = kFusedSmemDoubles * sizeof(double);
    VC_TRY(kernel_smem_optin(h));
    if (dp.fd == 6) fused_build_kernel<6><<<dp.n_frames, kFusedThreads, fsm, h->stream>>>(fa);
    else fused_build_kernel<9><<<dp.n_frames, kFusedThreads, fsm, h->stream>>>(fa);
    ++h->launches;
    n_vis_cost = dp.n_frames;
  } else {
    StageScope st(h, VCGPU_STAGE_BUILD);
    BuildArgs ba;
    ba.dp = vdp; ba.ctl = h->d_ctl; ba.which = which;
    ba.grp_start = h->d_grp_start; ba.grp_count = h->d_grp_count; ba.group_of = h->d_group_of;
    ba.r = h->d_r; ba.J = h->d_J; ba.n_obs = h->n_obs; ba.out[0] = h->blk[0]; ba.out[1] = h->blk[1]; ba.Cg = h->d_Cg;
    const size_t bsm = (2 * kBuildChunk * kMaxW + 9 * 9 + 9) * sizeof(double);
    if (dp.fd == 6) build_frames_kernel<6><<<dp.n_frames, kBuildThreads, bsm, h->stream>>>(ba);
    else build_frames_kernel<9><<<dp.n_frames, kBuildThreads, bsm, h->stream>>>(ba);
    ++h->launches;
  }
  if (dp.inertial) {  // after the fused launch: it is the one that writes the trial state
    StageScope st(h, VCGPU_STAGE_IMU_EVAL);
    VC_TRY(imu_evaluate(h, which, true, &n_imu_cost));
    VC_TRY(imu_accumulate(h, which));
  }
  const size_t NS = static_cast<size_t>(dp.G) * dp.G + dp.G;
  {
    StageScope st(h, VCGPU_STAGE_REDUCE);
    RedFinArgs ra;
    const bool multi = h->nranks > 1;
    const bool split = NS > 512;  // large global block: level 2 as its own parallel launches
    ra.dp = vdp; ra.ctl = h->d_ctl; ra.which = which; ra.decide_mode = multi ? -1 : decide_mode; ra.multi = multi ? 1 : 0;
    ra.level1_only = split ? 1 : 0;
    const bool sep = multi && dp.inertial;  // separator frames' gradients are normed after the all-reduce
    ra.gf_skip_below = sep ? dp.fd : 0;
    ra.gf_skip_from = sep ? static_cast<int64_t>(dp.n_own) * dp.fd : static_cast<int64_t>(dp.n_frames) * dp.fd;
    ra.Cg = h->d_Cg; ra.imuCg = dp.inertial ? imu_cg(h) : nullptr; ra.ni = dp.n_frames - 1; ra.imu_goff = dp.imu_goff;
    ra.imu_stride = kImuCgStride;
    ra.Cpart = h->d_Cpart; ra.red_part = h->d_red_part;
    ra.cost_part = h->d_cost_part; ra.n_cost_part = n_vis_cost;
    ra.imu_cost_part = imu_cost_part(h); ra.n_imu_cost_part = n_imu_cost;
    ra.step_part = with_step ? h->d_red : nullptr;
    // the last step_part slot is the globals' share: counted once (rank 0) in a sharded run
    ra.n_step_part = h->n_step_part - (h->rank == 0 ? 0 : 1);
    ra.n_frames_fd = dp.n_frames * dp.fd;
    ra.out[0] = h->blk[0]; ra.out[1] = h->blk[1]; ra.scalars = h->d_scalars; ra.counter = h->d_counter;
    VC_TRY(kernel_smem_optin(h));
    reduce_finalize_kernel<<<kReduceBlocks, 256, NS * sizeof(double), h->stream>>>(ra);
    ++h->launches;
    if (split) {
      sum_partials_sel_kernel<<<static_cast<int>((NS + 31) / 32), 256, 0, h->stream>>>(h->d_Cpart, kReduceBlocks, static_cast<int>(NS),
                                                                                       h->blk[0].C, h->blk[1].C, h->d_ctl, which);
      finalize_small_kernel<<<1, 256, 0, h->stream>>>(ra, kReduceBlocks);
      h->launches += 2;
    }
    if (multi) {  // sum the global blocks, cost and step scalars over the frame shards, then decide everywhere
      const int sep_fd = sep ? dp.fd : 0;
      mg_pack_kernel<<<8, 256, 0, h->stream>>>(dp.G, h->d_ctl, which, h->blk[0], h->blk[1], h->d_scalars, h->rank, h->nranks, h->d_mg,
                                               sep_fd, dp.n_frames, dp.ghost);
      ++h->launches;
      VC_TRY(all_reduce(h, h->d_mg, NS + 6 + h->nranks + 2 * static_cast<size_t>(sep_fd) * h->nranks));
      mg_unpack_decide_kernel<<<1, 256, 0, h->stream>>>(dp.G, h->d_ctl, which, h->blk[0], h->blk[1], h->d_scalars, h->nranks, h->d_mg,
                                                        decide_mode, sep_fd, h->d_sep);
      ++h->launches;
    }
  }
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

static int read_scalars(vcgpu_handle* h) {
  CUDA_TRY(h, cudaMemcpyAsync(h->h_scalars, h->d_scalars, kScCount * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  stage_collect(h);
  return VCGPU_OK;
}

// ------------------------------------------------------------------ damped arrow solve + state update
// Solves (H + D) step = -g on the blocks of the accepted buffer and writes the trial state into the
// other buffer.  D2x: explicit damping vector (inspection hook) or null for the LM rule.
static int solve_and_update(vcgpu_handle* h, const double* D2x, bool with_update_in_eval) {
  const DevProblem& dp = h->dp;
  const size_t NS = static_cast<size_t>(dp.G) * dp.G + dp.G;
  if (dp.inertial && imu_mega_applies(h)) {
    VC_TRY(imu_mega_solve(h, D2x, true));
  } else if (dp.inertial) {
    {
      StageScope st(h, VCGPU_STAGE_FRAME_SOLVE);
      VC_TRY(imu_chain_eliminate(h, D2x));
    }
    {
      StageScope st(h, VCGPU_STAGE_GLOBAL_SOLVE);
      VC_TRY(imu_chain_dense(h, D2x));
    }
    StageScope st(h, VCGPU_STAGE_BACKSUB);
    VC_TRY(imu_chain_backsub(h, D2x, h->materialize || !with_update_in_eval));
  } else {
    {
      StageScope st(h, VCGPU_STAGE_FRAME_SOLVE);
      SolveArgs sa;
      sa.dp = dp; sa.b[0] = h->blk[0]; sa.b[1] = h->blk[1]; sa.ctl = h->d_ctl; sa.scale = h->d_scale; sa.D2x = D2x;
      sa.X = h->d_X; sa.Spart = h->d_Spart; sa.scalars = h->d_scalars;
      const size_t ssm = (NS + static_cast<size_t>(kSolveWarps) * 2 * 6 * (dp.G + 1)) * sizeof(double);
      VC_TRY(kernel_smem_optin(h));
      frame_solve_kernel<6><<<h->n_solve_blocks, kSolveThreads, ssm, h->stream>>>(sa);
      ++h->launches;
    }
    const bool sum_in_solve = kSumInSolve && NS <= 1024 && h->nranks == 1;
    if (!sum_in_solve) {
      sum_partials_kernel<<<static_cast<int>((NS + 31) / 32), 256, 0, h->stream>>>(h->d_Spart, h->n_solve_blocks,
                                                                                   static_cast<int>(NS), h->d_Ssum, h->d_ctl);
      ++h->launches;
      VC_TRY(all_reduce(h, h->d_Ssum, NS));  // every rank then solves the same reduced system
    }
    {
      StageScope st(h, VCGPU_STAGE_GLOBAL_SOLVE);
      GlobalSolveArgs ga;
      ga.dp = dp; ga.b[0] = h->blk[0]; ga.b[1] = h->blk[1]; ga.ctl = h->d_ctl; ga.scale = h->d_scale; ga.D2x = D2x;
      ga.Ssum = h->d_Ssum; ga.Spart = h->d_Spart; ga.n_part = sum_in_solve ? h->n_solve_blocks : 0;
      ga.delta = h->d_delta; ga.scalars = h->d_scalars;
      VC_TRY(kernel_smem_optin(h));
      global_solve_kernel<<<1, 256, NS * sizeof(double), h->stream>>>(ga);
      ++h->launches;
    }
    if (h->materialize || !with_update_in_eval) {  // two-pass path / hook: explicit back-substitution kernel
      StageScope st(h, VCGPU_STAGE_BACKSUB);
      UpdateArgs ua;
      ua.dp = dp; ua.b[0] = h->blk[0]; ua.b[1] = h->blk[1]; ua.ctl = h->d_ctl; ua.scale = h->d_scale; ua.D2x = D2x;
      ua.X = h->d_X; ua.delta = h->d_delta; ua.state[0] = h->d_state[0]; ua.state[1] = h->d_state[1];
      ua.step_part = h->d_red; ua.sepdiag = nullptr;
      const int nb = (dp.n_frames + kUpdateWarps - 1) / kUpdateWarps;
      backsub_update_kernel<6><<<nb, 32 * kUpdateWarps, 0, h->stream>>>(ua);
      ++h->launches;
      h->n_step_part = nb + 1;
    }
  }
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

static double host_state_norm(const vcgpu_handle* h) {
  double s = 0;
  for (double v : h->h_T) s += v * v;
  if (h->flags.inertial) for (double v : h->h_v) s += v * v;
  for (int c = 0; c < h->n_cams; ++c) {
    for (int i = 0; i < 4; ++i) s += h->h_qck[4 * c + i] * h->h_qck[4 * c + i];
    for (int i = 0; i < 3; ++i) s += h->h_pck[3 * c + i] * h->h_pck[3 * c + i];
    for (int i = 0; i < num_intr(h->h_model[c]); ++i) s += h->h_intr[10 * c + i] * h->h_intr[10 * c + i];
  }
  if (h->flags.inertial) {
    s += h->h_g[0] * h->h_g[0] + h->h_g[1] * h->h_g[1] + h->h_ts * h->h_ts;
    for (int i = 0; i < 6; ++i) s += h->h_b[i] * h->h_b[i] + h->h_sf[i] * h->h_sf[i];
  }
  return std::sqrt(s);
}

static int num_residuals(const vcgpu_handle* h) {
  int64_t n = 0;
  if (h->flags.visual) {
    int64_t act = 0;
    for (uint8_t a : h->h_active) act += a;
    n += static_cast<int64_t>(2 * act * h->flags.visual_mult);
  }
  if (h->flags.inertial && h->n_frames > 1) n += static_cast<int64_t>(9 * (h->n_frames - 1) * h->flags.imu_mult);
  return static_cast<int>(n);
}

// ------------------------------------------------------------------ persistent vision kernel
static bool mega_applies(const vcgpu_handle* h) {
  return h->mega_warps > 0 && h->opts.strategy == 0 && !h->dp.inertial && (h->nranks == 1 || h->xchg_ready) && !h->materialize && !h->profiling && !h->multi_launch &&
         h->flags.visual && h->n_obs > 0;
}
// up to n_iters trust-region iterations in one cooperative launch (vc_mega.cuh)
static int mega_launch(vcgpu_handle* h, int n_iters) {
  const DevProblem& dp = h->dp;
  MegaArgs ma;
  ma.dp = dp; ma.ctl = h->d_ctl;
  ma.state[0] = h->d_state[0]; ma.state[1] = h->d_state[1];
  ma.blk[0] = h->blk[0]; ma.blk[1] = h->blk[1];
  ma.grp_start = h->d_grp_start; ma.grp_count = h->d_grp_count; ma.group_of = h->d_group_of;
  ma.pw = h->d_pw; ma.pc = h->d_pc; ma.mask = h->d_mask; ma.scale = h->d_scale; ma.X = h->d_X;
  ma.partS = h->d_partS; ma.partC = h->d_partC; ma.delta = h->d_delta; ma.scalars = h->d_scalars;
  ma.n_iters = n_iters; ma.n_warps = h->mega_warps;
  ma.rank = h->rank; ma.nranks = h->nranks;
  for (int r = 0; r < kMaxRanks; ++r) ma.xbuf[r] = reinterpret_cast<unsigned long long*>(h->xchg_peer[r]);
  ma.prof = h->phase_clocks ? h->d_prof : nullptr;
  void* args[] = {&ma};
  const size_t smem = mega_smem_doubles(dp.G, dp.n_cams, h->mega_warps) * sizeof(double);
  CUDA_TRY(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lm_mega_kernel), dim3(h->mega_grid), dim3(32 * h->mega_warps),
                                          args, smem, h->stream));
  ++h->launches;
  return VCGPU_OK;
}
// fold the persistent kernel's phase clocks into the stage times (call after a stream synchronise)
static int mega_collect_clocks(vcgpu_handle* h, int iters) {
  if (!h->phase_clocks || !h->d_prof) return VCGPU_OK;
  unsigned long long ns[kProfCount];
  CUDA_TRY(h, cudaMemcpy(ns, h->d_prof, sizeof ns, cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemset(h->d_prof, 0, sizeof ns));
  const int map[kProfCount] = {VCGPU_STAGE_FRAME_SOLVE, VCGPU_STAGE_GLOBAL_SOLVE, VCGPU_STAGE_BACKSUB, VCGPU_STAGE_BUILD,
                               VCGPU_STAGE_FINALIZE, VCGPU_STAGE_GRID_SYNC};
  for (int k = 0; k < kProfCount; ++k) { h->st_ms[map[k]] += ns[k] * 1e-6; h->st_n[map[k]] += iters; }
  return VCGPU_OK;
}

// ------------------------------------------------------------------ dogleg strategy (vc_dogleg.cuh)
static int dl_matvec(vcgpu_handle* h, const double* v, double* y, double* zpart) {
  const DevProblem& dp = h->dp;
  const int nb = (dp.n_frames + kMvWarps - 1) / kMvWarps;
  const size_t sm = static_cast<size_t>(kMvWarps) * dp.G * sizeof(double);
  if (dp.fd == 6) arrow_matvec_frames_kernel<6><<<nb, 32 * kMvWarps, sm, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, v, y, zpart);
  else arrow_matvec_frames_kernel<9><<<nb, 32 * kMvWarps, sm, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, v, y, zpart);
  ++h->launches;
  if (h->nranks > 1) {  // the globals' rows need E^T w_f of every rank's frames
    double* zsum = h->d_mg;
    arrow_matvec_globals_kernel<<<1, 256, 0, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, v, y, zpart, nb, 1, zsum);
    VC_TRY(all_reduce(h, zsum, dp.G));
    arrow_matvec_globals_kernel<<<1, 256, 0, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, v, y, zpart, nb, 2, zsum);
    h->launches += 2;
  } else {
    arrow_matvec_globals_kernel<<<1, 256, 0, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, v, y, zpart, nb, 0, nullptr);
    ++h->launches;
  }
  return VCGPU_OK;
}
// the three pairs of inner products of a dogleg iteration; frame shards sum them over the ranks (NCCL) first
static int dl_dots(vcgpu_handle* h, DlDotArgs da, int mode) {
  da.mode = mode;
  da.mg = h->nranks > 1 ? h->d_mg : nullptr;
  dl_dots_kernel<<<kDlBlocks, 256, 0, h->stream>>>(da);
  ++h->launches;
  if (h->nranks > 1) {
    VC_TRY(all_reduce(h, h->d_mg, 2));
    dl_dots_finish_kernel<<<1, 32, 0, h->stream>>>(h->d_ctl, mode, h->d_mg, h->d_scalars);
    ++h->launches;
  }
  return VCGPU_OK;
}
static int enqueue_dogleg_iteration(vcgpu_handle* h, bool weights) {
  const DevProblem& dp = h->dp;
  const int64_t np = static_cast<int64_t>(dp.n_frames) * dp.fd + dp.G;
  const int nmv = (dp.n_frames + kMvWarps - 1) / kMvWarps;
  VC_TRY(dev_alloc(h, &h->d_dl, 6 * static_cast<size_t>(np) + static_cast<size_t>(nmv) * dp.G));
  VC_TRY(dev_alloc(h, &h->d_dl_part, 4 * kDlBlocks));
  DlVecs v;
  v.diag = h->d_dl; v.grad = v.diag + np; v.gn = v.grad + np; v.vec = v.gn + np; v.Hv = v.vec + np; v.D2 = v.Hv + np;
  double* zpart = v.D2 + np;
  const int nblk = static_cast<int>((np + 255) / 256);
  DlDotArgs da;
  da.ctl = h->d_ctl; da.v = v; da.delta = h->d_delta; da.scalars = h->d_scalars; da.part = h->d_dl_part;
  da.counter = h->d_counter + 1; da.n = np;
  da.n_sum = (h->nranks > 1 && h->rank > 0) ? static_cast<int64_t>(dp.n_frames) * dp.fd : np;
  dl_prep_kernel<<<nblk, 256, 0, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale, v);
  ++h->launches;
  VC_TRY(dl_matvec(h, v.vec, v.Hv, zpart));
  VC_TRY(dl_dots(h, da, 0));
  VC_TRY(solve_and_update(h, v.D2, false));  // Gauss-Newton step with mu D^2 regularisation -> d_delta
  VC_TRY(dl_dots(h, da, 1));
  dl_combine_kernel<<<nblk, 256, 0, h->stream>>>(h->d_ctl, v, h->d_delta, np);
  ++h->launches;
  VC_TRY(dl_matvec(h, v.vec, v.Hv, zpart));
  VC_TRY(dl_dots(h, da, 2));
  {  // x (+) S step into the trial buffer, step statistics
    UpdateArgs ua;
    ua.dp = dp; ua.b[0] = h->blk[0]; ua.b[1] = h->blk[1]; ua.ctl = h->d_ctl; ua.scale = h->d_scale; ua.D2x = v.D2;
    ua.X = nullptr; ua.delta = h->d_delta; ua.state[0] = h->d_state[0]; ua.state[1] = h->d_state[1];
    ua.step_part = h->d_red; ua.sepdiag = nullptr;
    const int nb = (dp.n_frames + kUpdateWarps - 1) / kUpdateWarps;
    if (dp.fd == 6) backsub_update_kernel<6><<<nb, 32 * kUpdateWarps, 0, h->stream>>>(ua);
    else backsub_update_kernel<9><<<nb, 32 * kUpdateWarps, 0, h->stream>>>(ua);
    ++h->launches;
    h->n_step_part = nb + 1;
  }
  VC_TRY(evaluate_into(h, 1, true, -1));
  dl_decide_kernel<<<1, 32, 0, h->stream>>>(h->d_ctl, h->d_scalars);
  ++h->launches;
  if (weights) VC_TRY(imu_update_weights(h));
  CUDA_TRY(h, cudaGetLastError());
  return VCGPU_OK;
}

// one trust-region iteration, enqueued (no host wait)
static bool deferred_off() {
  static const bool off = std::getenv("VCGPU_NO_DEFERRED_WEIGHTS") != nullptr;  // A/B switch
  return off;
}
// the weight update of the last LM iteration of a persistent inertial run (earlier ones ride in the next solve launch)
static int imu_mega_finish_weights(vcgpu_handle* h, bool weights) {
  if (!weights || h->opts.strategy == 1 || !imu_mega_applies(h) || deferred_off()) return VCGPU_OK;
  return imu_update_weights(h, false, true);
}
static int enqueue_iteration(vcgpu_handle* h, bool weights) {
  if (h->opts.strategy == 1) return enqueue_dogleg_iteration(h, weights);
  if (mega_applies(h)) return mega_launch(h, 1);
  if (imu_mega_applies(h)) {
    // two cooperative launches: solve + update (+ the UpdateImuWeights the previous iteration's accepted step calls
    // for, on the CTAs the solve leaves idle), evaluate + decide.  The last iteration's update: imu_mega_finish_weights
    VC_TRY(imu_mega_solve(h, nullptr, true, weights && !deferred_off()));
    return imu_mega_eval(h, 1, true, 1, weights && deferred_off());
  }
  VC_TRY(solve_and_update(h, nullptr, false));
  VC_TRY(evaluate_into(h, 1, true, 1));
  if (weights) VC_TRY(imu_update_weights(h, true));  // the reference's iteration callback (vicalibrator.h:691)
  return VCGPU_OK;
}

// ------------------------------------------------------------------ the trust-region loop
static int run_solve(vcgpu_handle* h, vcgpu_iter_cb cb, void* user, vcgpu_summary* out, int fixed_iters) {
  VC_TRY(prepare(h));
  if (h->opts.strategy == 1 && h->nranks > 1 && h->flags.inertial)
    return fail(h, VCGPU_ERR_INVALID, "DOGLEG on frame shards covers the vision stages; sharded inertial solves use strategy 0 (LM)");
  const DevProblem& dp = h->dp;
  const vcgpu_options& o = h->opts;
  const int64_t np = static_cast<int64_t>(dp.n_frames) * dp.fd + dp.G;
  const long launches0 = h->launches;
  const int max_it = fixed_iters > 0 ? fixed_iters : o.max_iters;
  const bool weights = o.update_imu_weights && dp.inertial;
  VC_TRY(ctl_reset(h, fixed_iters > 0 ? 1 : 0, max_it));
  h->h_ctl->x_norm = host_state_norm(h);
  VC_TRY(ctl_upload(h));
  if (weights) VC_TRY(imu_update_weights(h));  // vicalibrator.h:955
  VC_TRY(evaluate_into(h, 0, false, 0));
  if (o.jacobi_scaling) {
    jacobi_scale_kernel<<<static_cast<int>((np + 255) / 256), 256, 0, h->stream>>>(dp, h->blk[0], h->blk[1], h->d_ctl, h->d_scale,
                                                                                   (h->nranks > 1 && dp.inertial) ? h->d_sep : nullptr);
    ++h->launches;
  } else {
    std::vector<double> ones(np, 1.0);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_scale, ones.data(), np * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  if (weights) VC_TRY(imu_update_weights(h));  // callback after iteration 0 (vicalibrator.h:691)
  const bool per_iteration = cb != nullptr || o.update_state_every_iteration || h->flush_l2 || h->profiling;
  auto report = [&](void) -> int {  // invoke the user's callback with the last iteration's summary
    const Ctl& c = *h->h_ctl;
    vcgpu_iteration it;
    it.iteration = c.iter; it.step_is_successful = c.last_accepted; it.cost = c.cost; it.cost_change = c.last_cost_change;
    it.gradient_max_norm = c.gmax; it.gradient_norm = c.last_accepted ? c.gnorm : 0.0; it.step_norm = c.last_step_norm;
    it.relative_decrease = c.last_rho; it.trust_region_radius = c.radius;
    if (o.update_state_every_iteration) { VC_TRY(download_state(h)); write_mirrors(h); }
    if (cb && cb(&it, user)) {
      if (!c.done) {
        h->h_ctl->done = 1 + VCGPU_TERM_CALLBACK;
        CUDA_TRY(h, cudaMemcpyAsync(&h->d_ctl->done, &h->h_ctl->done, sizeof(int), cudaMemcpyHostToDevice, h->stream));
      }
    }
    return VCGPU_OK;
  };
  if (per_iteration && fixed_iters <= 0) {
    VC_TRY(ctl_download(h));
    VC_TRY(report());
  }
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  double flushed_ms = 0.0;
  if (per_iteration) {
    for (int k = 0; k < max_it && !h->h_ctl->done; ++k) {
      if (h->flush_l2) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_flush, k & 0xff, 256u << 20, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->it_ev[0], h->stream));
      }
      VC_TRY(enqueue_iteration(h, weights));
      if (h->flush_l2) {
        CUDA_TRY(h, cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->it_ev[1], h->stream));
      }
      VC_TRY(ctl_download(h));
      if (h->flush_l2) {
        float ims = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&ims, h->it_ev[0], h->it_ev[1]));
        flushed_ms += ims;
      }
      if (fixed_iters <= 0) VC_TRY(report());
    }
  } else {
    // free-running: enqueue batches; the device decides, the host only looks at `done` between batches
    const int batch = fixed_iters > 0 ? 32 : 4;
    int queued = 0;
    if (mega_applies(h)) {  // the whole loop in one launch; the device stops at convergence
      VC_TRY(mega_launch(h, max_it));
      queued = max_it;
    }
    while (queued < max_it) {
      const int n = std::min(batch, max_it - queued);
      for (int k = 0; k < n; ++k) VC_TRY(enqueue_iteration(h, weights));
      queued += n;
      if (fixed_iters > 0 && queued < max_it) continue;  // benchmark mode: no intermediate sync at all
      VC_TRY(ctl_download(h));
      if (h->h_ctl->done) break;
    }
  }
  VC_TRY(imu_mega_finish_weights(h, weights));
  VC_TRY(wts_join(h));
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  VC_TRY(ctl_download(h));
  CUDA_TRY(h, cudaEventSynchronize(h->ev1));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  const Ctl& c = *h->h_ctl;
  if (c.done == kMegaCommFailed) return fail(h, VCGPU_ERR_COMM, "a peer GPU never reached the in-kernel exchange (2 s timeout)");
  vcgpu_summary sum;
  std::memset(&sum, 0, sizeof sum);
  sum.num_residuals = num_residuals(h);
  sum.iterations = c.iter;
  sum.successful_steps = c.successful;
  sum.termination = c.done ? c.done - 1 : VCGPU_TERM_NO_CONVERGENCE;
  sum.initial_cost = c.initial_cost;
  sum.final_cost = c.cost;
  sum.device_seconds = (h->flush_l2 ? flushed_ms : ms) * 1e-3;
  sum.kernel_launches = static_cast<int>(h->launches - launches0);
  VC_TRY(mega_collect_clocks(h, c.iter));
  VC_TRY(imu_mega_collect_clocks(h, c.iter));
  h->blocks_valid = true;
  VC_TRY(download_state(h));
  write_mirrors(h);
  if (out) *out = sum;
  return VCGPU_OK;
}
