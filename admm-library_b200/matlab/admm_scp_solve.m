function [x, z, u, hist] = admm_scp_solve(prob, opts, scp)
%ADMM_SCP_SOLVE  Sequential convex programming on NVIDIA B200: rendezvous with nonlinear relative dynamics,
% re-linearised per problem ON THE GPU each pass, convex subproblems on the batched ADMM kernels (MEX -> C ABI ->
% CUDA, admmb_scp_solve).  Same calling surface as the oracle oracle/admm_scp.m.
%
%   [x, z, u, hist] = admm_scp_solve(prob, opts, scp)
%
%   prob.N, prob.s0 [6 x Bsz], prob.block_type int32 [3N+2], prob.block_par [8 x (3N+2) x Bp];
%   optional prob.q [n x Bq], per-problem prob.Q [6 x 6 x (N+1) x Bsz], prob.R [3 x 3 x N x Bsz]   (no A, B, c)
%   opts: as admm_solve; adapt_rho and history must be off, xupdate 'auto' or 'riccati', precision 'fp64'
%   scp : model ('nl_circular' | 'nl_elliptic': chief on a Kepler orbit, then scp.e, scp.theta0 [Bsz], R0 = semi-major axis), T, R0, nmm (default 1), substeps (default 8), max_pass, tol_abs, tol_rel,
%         control ('zoh' (default): thrust acceleration held over a stage | 'impulsive': velocity increment, then a coast)
%
%   hist: the fields of admm_solve for every problem's LAST convex solve, plus
%         scp_passes, scp_status (0 trajectory converged, 1 max_pass), scp_step [Bsz], scp_iters_total int64 [Bsz],
%         scp_hist_step [max_pass x Bsz] (NaN after a problem's exit),
%         scp_stats = [converged, sum of ADMM iterations over all passes, max passes, sum of passes].

    if nargin < 3, error('admm:ADMMB_E_BADARG', 'usage: admm_scp_solve(prob, opts, scp)'); end
    d = struct('rho', 1.0, 'alpha', 1.0, 'abstol', 1e-6, 'reltol', 1e-6, 'max_iter', 1000, ...
               'adapt_rho', 0, 'xupdate', 'auto', 'precision', 'fp64', 'history', 0, 'gpus', 0, 'chunk', 0, 'kernel', 0);
    f = fieldnames(d);
    for i = 1:numel(f)
        if ~isfield(opts, f{i}), opts.(f{i}) = d.(f{i}); end
    end
    prob.block_type = int32(prob.block_type);
    [x, z, u, hist] = admm_mex(prob, opts, scp);   % third argument: the gateway calls admmb_scp_solve
end
