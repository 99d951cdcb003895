function [x, z, u, hist] = admm_ocp(prob, opts)
%ADMM_OCP  ORACLE (test infrastructure, not product code): minimal MATLAB ADMM for the batched
% convex optimal-control QPs of this repository, written to the one algorithmic statement the
% reference makes (/root/reference/README.md:2 "Implementation of Alternating Direction Method of
% Multipliers for astrodynamics problems").  BASELINE.json's north_star mandates this file because
% the mounted reference tree carries only README/LICENSE.
%
% PARITY STATUS: unpinned by the reference (it ships no code, tests or golden vectors), and this
% file is UNEXECUTED in this environment: neither MATLAB nor Octave is installed.  Its executable
% restatements are oracle/admm_ocp.py (NumPy, function for function) and oracle/admm_ocp_cpu.c
% (canonical operation order); parity claims are made against those.
%
%   [x, z, u, hist] = admm_ocp(prob, opts)         same surface as matlab/admm_solve.m
%
%   minimise 1/2 x'Px + q'x + g(z_J)   s.t.  Gx = h,  x_J = z_J          (scaled-form ADMM, Boyd 2011)
%   x = (s_0,a_0,...,s_{N-1},a_{N-1},s_N), dynamics s_{k+1} = A_k s_k + B_k a_k + c_k, s_0 = s_init,
%   P = blkdiag(Q_0,R_0,...,Q_N), g separable over consecutive 3-vectors ("blocks"), J = split blocks
%   (every block whose type is not 8 = NONE).
%
%   iteration:  x  = argmin 1/2 x'(P/rho)x + (q/rho)'x + 1/2|x_J - (z-u)_J|^2  s.t. Gx = h   (a2)
%               xh = alpha x + (1-alpha) z ;  z = prox_{g/rho}(xh + u) ;  u = u + xh - z      (a3,a4)
%               r = |x_J - z_J| ; s = rho |z - z_old| ; stop when r < eps_pri and s < eps_dual
%               every `adapt_every` iterations (until `adapt_until`): residual balancing of rho (a5)

    N = size(prob.A, 3);  n = 9*N + 6;  nb = 3*N + 2;
    Bsz = size(prob.s0, 2);  Bd = size(prob.A, 4);
    o = defaults(opts);
    bt = double(prob.block_type(:));
    wblk = double(bt ~= 8);  w = kron(wblk, ones(3,1));  nsplit = sum(w);
    has_c = isfield(prob,'c') && ~isempty(prob.c);
    has_Q = isfield(prob,'Q') && ~isempty(prob.Q);
    has_R = isfield(prob,'R') && ~isempty(prob.R);
    has_q = isfield(prob,'q') && ~isempty(prob.q);
    x = zeros(n,Bsz); z = zeros(n,Bsz); u = zeros(n,Bsz);
    hist.iters = zeros(Bsz,1,'int32'); hist.status = ones(Bsz,1,'int32');
    hist.r_norm = zeros(Bsz,1); hist.s_norm = zeros(Bsz,1); hist.eps_pri = zeros(Bsz,1);
    hist.eps_dual = zeros(Bsz,1); hist.rho = zeros(Bsz,1);
    if o.history
        hist.hist_r_norm = nan(o.max_iter,Bsz); hist.hist_s_norm = nan(o.max_iter,Bsz);
        hist.hist_eps_pri = nan(o.max_iter,Bsz); hist.hist_eps_dual = nan(o.max_iter,Bsz);
        hist.hist_rho = nan(o.max_iter,Bsz);
    end
    nrefac = 0;
    for p = 1:Bsz                                   % problems are independent
        pd = min(p, Bd);
        A = prob.A(:,:,:,pd);  B = prob.B(:,:,:,pd);
        c = []; Q = []; R = []; q = [];
        if has_c, c = prob.c(:,:,pd); end
        if has_Q, Q = prob.Q(:,:,:,pd); end
        if has_R, R = prob.R(:,:,:,pd); end
        if has_q, q = prob.q(:, min(p, size(prob.q,2))); end
        par = prob.block_par(:,:, min(p, size(prob.block_par,3)));
        s0 = prob.s0(:,p);
        rho = o.rho;  if isfield(prob,'rho0') && ~isempty(prob.rho0), rho = prob.rho0(p); end
        zp = zeros(n,1); up = zeros(n,1);
        if isfield(prob,'z0') && ~isempty(prob.z0), zp = w .* prob.z0(:,p); end
        if isfield(prob,'u0') && ~isempty(prob.u0), up = w .* prob.u0(:,p); end
        fac = riccati_factor(A, B, c, Q, R, rho, wblk);                     % a1, once per rho
        if strcmp(o.xupdate,'dense'), dfac = kkt_dense_factor(A, B, c, Q, R, rho, w); end   % a1'
        status = 1;  k = 0;  xp = zeros(n,1);
        for k = 1:o.max_iter
            rt = w .* (zp - up);
            if has_q, rt = rt - q / rho; end
            if strcmp(o.xupdate,'dense')
                xp = dfac.M * rt + dfac.S * s0 + dfac.mc;                   % a2'
            else
                xp = xupdate_riccati(fac, A, B, c, s0, rt);                 % a2
            end
            xh = o.alpha * xp + (1 - o.alpha) * zp;
            zn = w .* prox_blocks(xh + up, bt, par, 1/rho);                 % a3
            un = w .* (xh + up - zn);                                       % a4
            r_norm = norm(w .* (xp - zn));  s_norm = rho * norm(zn - zp);
            eps_pri  = sqrt(nsplit)*o.abstol + o.reltol * max(norm(w .* xp), norm(zn));
            eps_dual = sqrt(nsplit)*o.abstol + o.reltol * rho * norm(un);
            zp = zn;  up = un;
            if o.history
                hist.hist_r_norm(k,p) = r_norm; hist.hist_s_norm(k,p) = s_norm;
                hist.hist_eps_pri(k,p) = eps_pri; hist.hist_eps_dual(k,p) = eps_dual; hist.hist_rho(k,p) = rho;
            end
            if ~(isfinite(r_norm) && isfinite(s_norm)), status = 2; break; end
            if r_norm < eps_pri && s_norm < eps_dual, status = 0; break; end
            if o.adapt_rho && mod(k, o.adapt_every) == 0 && k < o.max_iter && (o.adapt_until <= 0 || k <= o.adapt_until)
                [rho_new, usc] = adapt_rho(r_norm, s_norm, rho, o.adapt_mu, o.adapt_tau);   % a5
                if rho_new ~= rho
                    rho = rho_new;  up = up * usc;
                    if has_Q || has_R
                        fac = riccati_factor(A, B, c, Q, R, rho, wblk);  nrefac = nrefac + 1;
                    end
                end
            end
        end
        x(:,p) = xp;  z(:,p) = zp + (1 - w) .* xp;  u(:,p) = up;
        hist.iters(p) = k; hist.status(p) = status; hist.rho(p) = rho;
        hist.r_norm(p) = r_norm; hist.s_norm(p) = s_norm; hist.eps_pri(p) = eps_pri; hist.eps_dual(p) = eps_dual;
    end
    hist.stats = [sum(hist.status == 0); sum(double(hist.iters)); max(double(hist.iters)); nrefac];
end

function o = defaults(opts)
    o = struct('rho',1,'alpha',1,'abstol',1e-6,'reltol',1e-6,'max_iter',1000,'adapt_rho',0,'adapt_mu',10, ...
               'adapt_tau',2,'adapt_every',25,'adapt_until',0,'xupdate','auto','history',0);
    f = fieldnames(opts);
    for i = 1:numel(f), o.(f{i}) = opts.(f{i}); end
end

% ---- a1 ------------------------------------------------------------------------------------
function fac = riccati_factor(A, B, c, Q, R, rho, wblk)
    N = size(A,3);
    Ds = @(k) diag(kron(wblk(3*k+1:3*k+2), ones(3,1)));      % k = 0..N (0-based stage)
    Da = @(k) wblk(3*k+3) * eye(3);
    P = Ds(N);  if ~isempty(Q), P = P + Q(:,:,N+1)/rho; end
    fac.K = zeros(3,6,N); fac.E = zeros(3,6,N); fac.Hinv = zeros(3,3,N); fac.Acl = zeros(6,6,N); fac.chat = zeros(6,N);
    for k = N-1:-1:0
        Ak = A(:,:,k+1);  Bk = B(:,:,k+1);
        Wa = Da(k);  if ~isempty(R), Wa = Wa + R(:,:,k+1)/rho; end
        H = Wa + Bk' * P * Bk;
        Hi = inv(H);  Hi = (Hi + Hi')/2;
        Kk = -Hi * (Bk' * P * Ak);
        fac.K(:,:,k+1) = Kk;  fac.E(:,:,k+1) = Hi * Bk';  fac.Hinv(:,:,k+1) = Hi;
        fac.Acl(:,:,k+1) = Ak + Bk * Kk;
        if ~isempty(c), fac.chat(:,k+1) = P * c(:,k+1); end
        Ws = Ds(k);  if ~isempty(Q), Ws = Ws + Q(:,:,k+1)/rho; end
        P = Ws + Ak' * P * fac.Acl(:,:,k+1);  P = (P + P')/2;
    end
end

% ---- a2 ------------------------------------------------------------------------------------
function x = xupdate_riccati(fac, A, B, c, s0, rt)
    N = size(A,3);  x = zeros(9*N+6,1);  d = zeros(3,N);
    g = rt(9*N+1:9*N+6);
    for k = N-1:-1:0
        if ~isempty(c), g = g - fac.chat(:,k+1); end
        rs = rt(9*k+1:9*k+6);  ra = rt(9*k+7:9*k+9);
        d(:,k+1) = fac.Hinv(:,:,k+1) * ra + fac.E(:,:,k+1) * g;
        g = rs + fac.K(:,:,k+1)' * ra + fac.Acl(:,:,k+1)' * g;
    end
    s = s0;
    for k = 0:N-1
        a = d(:,k+1) + fac.K(:,:,k+1) * s;
        x(9*k+1:9*k+6) = s;  x(9*k+7:9*k+9) = a;
        s = A(:,:,k+1) * s + B(:,:,k+1) * a;
        if ~isempty(c), s = s + c(:,k+1); end
    end
    x(9*N+1:9*N+6) = s;
end

% ---- a1' -----------------------------------------------------------------------------------
function dfac = kkt_dense_factor(A, B, c, Q, R, rho, w)
    N = size(A,3);  n = 9*N+6;  m = 6*(N+1);
    G = zeros(m,n);  G(1:6,1:6) = eye(6);  Pm = zeros(n);
    for k = 0:N-1
        r = 6*(k+1);
        G(r+1:r+6, 9*k+1:9*k+6) = -A(:,:,k+1);  G(r+1:r+6, 9*k+7:9*k+9) = -B(:,:,k+1);
        G(r+1:r+6, 9*(k+1)+1:9*(k+1)+6) = eye(6);
    end
    for k = 0:N
        if ~isempty(Q), Pm(9*k+1:9*k+6, 9*k+1:9*k+6) = Q(:,:,k+1); end
        if ~isempty(R) && k < N, Pm(9*k+7:9*k+9, 9*k+7:9*k+9) = R(:,:,k+1); end
    end
    Ki = inv([diag(w) + Pm/rho, G'; G, zeros(m)]);
    dfac.M = Ki(1:n,1:n);  dfac.S = Ki(1:n,n+1:n+6);
    if isempty(c), dfac.mc = zeros(n,1); else, dfac.mc = Ki(1:n,n+7:end) * c(:); end
end

% ---- a3 ------------------------------------------------------------------------------------
function z = prox_blocks(v, bt, par, rinv)
    z = v;
    for b = 1:numel(bt)
        i = 3*(b-1) + (1:3);  w = v(i);  p = par(:,b);
        kap = p(1) * rinv;  rad = p(2);  lo = p(3:5);  hi = p(6:8);
        switch bt(b)
            case {0, 8}, z(i) = w;
            case 1, z(i) = sign(w) .* max(abs(w) - kap, 0);
            case 2, z(i) = min(max(sign(w) .* max(abs(w) - kap, 0), lo), hi);
            case {3, 4}
                nrm = norm(w);
                if nrm > kap
                    mag = nrm - kap;  if bt(b) == 4, mag = min(mag, rad); end
                    z(i) = (mag / nrm) * w;
                else
                    z(i) = 0;
                end
            case 5, z(i) = min(max(w, lo), hi);
            case 6
                dv = w - lo;  nrm = norm(dv);
                if nrm > rad, z(i) = lo + (rad / nrm) * dv; end
            case 7, z(i) = lo;
        end
    end
end

% ---- a5 ------------------------------------------------------------------------------------
function [rho, usc] = adapt_rho(r_norm, s_norm, rho, mu, tau)
    usc = 1;
    if r_norm > mu * s_norm
        if rho * tau <= 1e6, rho = rho * tau;  usc = 1/tau; end
    elseif s_norm > mu * r_norm
        if rho / tau >= 1e-6, rho = rho / tau;  usc = tau; end
    end
end
