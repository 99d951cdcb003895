function [x, z, u, hist] = admm_solve(prob, opts)
%ADMM_SOLVE  Batched ADMM for convex optimal-control QPs on NVIDIA B200 (MEX -> C ABI -> CUDA).
%
%   [x, z, u, hist] = admm_solve(prob, opts)
%
% Same calling surface as the oracle oracle/admm_ocp.m: problem struct in; primal iterate x, split
% copy z, scaled dual u and residual history out; rho / alpha / abstol / reltol with the semantics of
% Boyd et al. (2011).  No gpuArray is involved and there is no CPU fallback: the call fails with
% admm:ADMMB_E_NODEVICE when no CUDA device is usable.
%
%   prob.A  [6 x 6 x N x Bd]   prob.B  [6 x 3 x N x Bd]   prob.c  [6 x N x Bd]     (Bd = 1 or Bsz)
%   prob.Q  [6 x 6 x (N+1) x Bd]   prob.R [3 x 3 x N x Bd]   prob.q [n x Bq]       (optional)
%   prob.s0 [6 x Bsz]          prob.block_type int32 [3N+2]   prob.block_par [8 x (3N+2) x Bp]
%   prob.z0, prob.u0 [n x Bsz], prob.rho0 [Bsz]                                    (optional warm start)
%   opts: rho alpha abstol reltol max_iter adapt_rho adapt_mu adapt_tau adapt_every adapt_until
%         xupdate ('auto'|'dense'|'riccati') precision ('fp64'|'tf32') history gpus chunk
%         kernel (0 = auto; ADMMB_KERNEL_* code, tests only) tf32_switch tf32_refresh (0 = defaults)
%         precision 'fp64' (default): results bit-identical to the oracle admm_ocp.m order of operations.
%         precision 'tf32': the tensor cores may be used (x-update increments as TF32x3 GEMMs, FP64 accumulation) --
%         with xupdate 'dense' throughout, with 'auto' once few problems are still running; results then follow the
%         FP64 iteration within a tolerance (same converged set down to 1e-8, iterates not bit-identical).
%
%   x, z, u  [n x Bsz],  n = 9N+6, stage-interleaved (s_0,a_0,...,s_N)
%   hist.iters, hist.status (0 converged, 1 max_iter, 2 nan) [Bsz]; hist.r_norm, s_norm, eps_pri,
%   eps_dual, rho [Bsz] (finals) and, with opts.history, hist.hist_* [max_iter x Bsz];
%   hist.stats = [converged, sum iters, max iters, refactorisations].
%
% The reference repository (SergioCdV/ADMM-library as mounted: README.md:1-2, LICENSE) defines no
% calling surface of its own; this one is fixed by BASELINE.json's north_star.

    if nargin < 2, opts = struct(); end
    d = struct('rho', 1.0, 'alpha', 1.0, 'abstol', 1e-6, 'reltol', 1e-6, 'max_iter', 1000, ...
               'adapt_rho', 0, 'adapt_mu', 10.0, 'adapt_tau', 2.0, 'adapt_every', 25, 'adapt_until', 0, ...
               'xupdate', 'auto', 'precision', 'fp64', 'history', 0, 'gpus', 0, 'chunk', 0, ...
               'kernel', 0, 'tf32_switch', 0, 'tf32_refresh', 0);
    f = fieldnames(d);
    for i = 1:numel(f)
        if ~isfield(opts, f{i}), opts.(f{i}) = d.(f{i}); end
    end
    prob.block_type = int32(prob.block_type);
    [x, z, u, hist] = admm_mex(prob, opts);     % thin gateway: validates fields, passes raw pointers
end
