function [x, z, u, hist] = admm_scp(prob, opts, scp)
%ADMM_SCP  ORACLE (test infrastructure, not product code): sequential convex programming around admm_ocp.m
% for rendezvous with NONLINEAR relative dynamics (SURVEY.md 8(f-4)).  Same surface as matlab/admm_scp_solve.m.
%
% PARITY STATUS: unpinned by the reference (/root/reference/README.md:1-2 is its whole content: no SCP loop, no
% code, no vectors) and UNEXECUTED here (no MATLAB / Octave in the image).  The executable restatement, in the same
% operation order, is oracle/scp_ocp.py; parity claims are made against that file.
%
%   prob: N, s0 [6 x Bsz], block_type, block_par, optional q / per-problem Q, R  (A, B, c are produced here)
%   scp : T (stage length), R0 (radius of the chief's circular orbit), nmm (mean motion, default 1),
%         substeps (RK4 steps per stage, default 8), max_pass, tol_abs, tol_rel,
%         control ('zoh' thrust acceleration held over the stage (default) | 'impulsive' velocity increment + coast)
%
%   model: deputy about a chief on a circular orbit, LVLH frame (x radial, y along-track, z cross-track),
%          mu = n^2 R0^3, full two-body gravity, zero-order-hold thrust acceleration a:
%            rho = [R0 + x; y; z],  g = n^2 - mu/|rho|^3,  k = mu/|rho|^3
%            x'' =  2 n y' + g (R0 + x) + a_x ,   y'' = -2 n x' + g y + a_y ,   z'' = -k z + a_z
%          (the elliptic-chief model 'nl_elliptic' of the library is stated in oracle/scp_ocp.py only: this text covers
%          'nl_circular' with both control models)
%   pass : A_k = dF/ds, B_k = dF/da, c_k = F(s_k, a_k) - A_k s_k - B_k a_k about the reference (RK4 of the state and
%          its variational equations); convex subproblem by admm_ocp, warm-started from the previous pass;
%          a problem leaves the loop when max|x - x_ref| <= tol_abs + tol_rel max|x|.
%          First reference: free drift from s0.

    if isfield(scp,'model') && ~strcmp(scp.model,'nl_circular'), error('admm_scp: only the nl_circular model is written out here'); end
    N = prob.N;  n = 9*N + 6;  Bsz = size(prob.s0, 2);
    if ~isfield(scp,'nmm') || scp.nmm == 0, scp.nmm = 1; end
    if ~isfield(scp,'substeps') || scp.substeps == 0, scp.substeps = 8; end
    xref = zeros(n, Bsz);
    A = zeros(6,6,N,Bsz); B = zeros(6,3,N,Bsz); c = zeros(6,N,Bsz);
    s = prob.s0;
    for k = 1:N                                      % free drift, linearised on the way
        xref(9*(k-1)+(1:6), :) = s;
        [s, A(:,:,k,:), B(:,:,k,:), c(:,k,:)] = stage(s, zeros(3,Bsz), scp);
    end
    xref(9*N+(1:6), :) = s;
    x = zeros(n,Bsz); z = x; u = x;
    hist.scp_passes = zeros(Bsz,1,'int32'); hist.scp_status = ones(Bsz,1,'int32');
    hist.scp_step = nan(Bsz,1); hist.scp_iters_total = zeros(Bsz,1,'int64');
    hist.scp_hist_step = nan(scp.max_pass, Bsz);
    hist.iters = zeros(Bsz,1,'int32'); hist.status = zeros(Bsz,1,'int32');
    active = true(Bsz,1);
    for p = 1:scp.max_pass
        idx = find(active);
        if isempty(idx), break; end
        if p > 1
            for k = 1:N
                [~, A(:,:,k,idx), B(:,:,k,idx), c(:,k,idx)] = ...
                    stage(xref(9*(k-1)+(1:6), idx), xref(9*(k-1)+(7:9), idx), scp);
            end
        end
        sub = prob;  sub.A = A(:,:,:,idx); sub.B = B(:,:,:,idx); sub.c = c(:,:,idx); sub.s0 = prob.s0(:,idx);
        if isfield(prob,'Q') && ~isempty(prob.Q), sub.Q = prob.Q(:,:,:,idx); end
        if isfield(prob,'R') && ~isempty(prob.R), sub.R = prob.R(:,:,:,idx); end
        if isfield(prob,'q') && size(prob.q,2) > 1, sub.q = prob.q(:,idx); end
        if size(prob.block_par,3) > 1, sub.block_par = prob.block_par(:,:,idx); end
        if p > 1, sub.z0 = z(:,idx); sub.u0 = u(:,idx); end
        [xs, zs, us, h] = admm_ocp(sub, opts);
        x(:,idx) = xs; z(:,idx) = zs; u(:,idx) = us;
        hist.iters(idx) = h.iters; hist.status(idx) = h.status;
        hist.scp_iters_total(idx) = hist.scp_iters_total(idx) + int64(h.iters);
        step = max(abs(xs - xref(:,idx)), [], 1, 'includenan').';
        scale = max(abs(xs), [], 1, 'includenan').';
        xref(:,idx) = xs;
        hist.scp_passes(idx) = p; hist.scp_step(idx) = step; hist.scp_hist_step(p, idx) = step;
        done = step <= scp.tol_abs + scp.tol_rel * scale;
        hist.scp_status(idx(done)) = 0;
        active(idx(done)) = false;
    end
end

function [F, A, B, c] = stage(sr, ar, scp)
% RK4 of the state and, column by column, of [Phi | Gamma]; sr [6 x B], ar [3 x B]
    if isfield(scp, 'control') && strcmp(scp.control, 'impulsive')
        % velocity increment at the start of the stage, then a coast: A = dF/ds at the post-impulse state, B = A(:,4:6)
        sin = sr;  sin(4:6,:) = sr(4:6,:) + ar;
        coast = scp;  coast.control = 'zoh';
        [F, A, ~, ~] = stage(sin, zeros(3, size(sr,2)), coast);
        B = A(:,4:6,1,:);
        c = F;
        for j = 1:6
            y = squeeze(A(:,j,1,:));  if size(sr,2) == 1, y = y(:); end
            c = c - y .* sr(j,:);
            if j > 3, c = c - y .* ar(j-3,:); end
        end
        return;
    end
    Bsz = size(sr, 2);  n = scp.nmm;  R0 = scp.R0;  n2 = n*n;  tn = 2*n;
    dt = scp.T / scp.substeps;  hdt = 0.5*dt;  dt6 = dt/6;
    A = zeros(6,6,1,Bsz); B = zeros(6,3,1,Bsz);
    s = sr;  Y = zeros(6, 9, Bsz);
    for j = 1:6, Y(j,j,:) = 1; end
    for ss = 1:scp.substeps
        c1 = coeffs(s, R0, n2);          k1 = fstate(c1, s, ar, tn);
        s2 = s + hdt*k1;  c2 = coeffs(s2, R0, n2);  k2 = fstate(c2, s2, ar, tn);
        s3 = s + hdt*k2;  c3 = coeffs(s3, R0, n2);  k3 = fstate(c3, s3, ar, tn);
        s4 = s + dt*k3;   c4 = coeffs(s4, R0, n2);  k4 = fstate(c4, s4, ar, tn);
        for j = 1:9
            fr = 0; if j > 6, fr = j - 3; end          % Gamma column j-6 is forced on row 3 + (j-6)
            y = squeeze(Y(:,j,:));  if Bsz == 1, y = y(:); end
            l1 = fcol(c1, s,  y,          tn, fr);
            l2 = fcol(c2, s2, y + hdt*l1, tn, fr);
            l3 = fcol(c3, s3, y + hdt*l2, tn, fr);
            l4 = fcol(c4, s4, y + dt*l3,  tn, fr);
            Y(:,j,:) = y + dt6 * (((l1 + 2*l2) + 2*l3) + l4);
        end
        s = s + dt6 * (((k1 + 2*k2) + 2*k3) + k4);
    end
    F = s;  c = s;
    for j = 1:9
        y = squeeze(Y(:,j,:));  if Bsz == 1, y = y(:); end
        if j <= 6, A(:,j,1,:) = y; v = sr(j,:); else, B(:,j-6,1,:) = y; v = ar(j-6,:); end
        c = c - y .* v;
    end
end

function cf = coeffs(s, R0, n2)
    cf.rx = R0 + s(1,:);
    q = ((((2*R0)*s(1,:) + s(1,:).*s(1,:)) + s(2,:).*s(2,:)) + s(3,:).*s(3,:)) / (R0*R0);
    w = 1 + q;  w32 = w .* sqrt(w);
    cf.k = n2 ./ w32;
    cf.g = (n2 * (q .* ((3 + 3*q) + q.*q))) ./ (w32 .* (1 + w32));     % n^2 - mu/d^3 without cancellation
    cf.m = (3*cf.k) ./ ((R0*R0) * w);
end

function ds = fstate(cf, s, a, tn)
    ds = [s(4,:); s(5,:); s(6,:);
          (tn*s(5,:) + cf.g.*cf.rx) + a(1,:);
          ((-tn)*s(4,:) + cf.g.*s(2,:)) + a(2,:);
          (-cf.k).*s(3,:) + a(3,:)];
end

function dy = fcol(cf, s, y, tn, fr)
    j30 = cf.g + cf.m.*(cf.rx.*cf.rx);  j31 = cf.m.*(cf.rx.*s(2,:));  j32 = cf.m.*(cf.rx.*s(3,:));
    j41 = cf.g + cf.m.*(s(2,:).*s(2,:)); j42 = cf.m.*(s(2,:).*s(3,:)); j52 = (-cf.k) + cf.m.*(s(3,:).*s(3,:));
    dy = [y(4,:); y(5,:); y(6,:);
          ((j30.*y(1,:) + j31.*y(2,:)) + j32.*y(3,:)) + tn*y(5,:);
          ((j31.*y(1,:) + j41.*y(2,:)) + j42.*y(3,:)) + (-tn)*y(4,:);
          (j32.*y(1,:) + j42.*y(2,:)) + j52.*y(3,:)];
    if fr > 0, dy(fr+1,:) = dy(fr+1,:) + 1; end
end
