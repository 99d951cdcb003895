% CHECK_OCTAVE  ORACLE (test infrastructure): runs the MATLAB-text oracles oracle/admm_ocp.m and oracle/admm_scp.m under
% GNU Octave (>= 7, for jsondecode) or MATLAB on two committed fixtures and compares them with the values the executable
% restatements (oracle/admm_ocp_cpu.c, oracle/scp_ocp.py) produced -- SURVEY.md 8(f-4), second half: "a real
% MATLAB/Octave CI job that executes admm_ocp.m".
%
% STATUS: UNEXECUTED in the build image (neither MATLAB nor Octave is installed there); provided for
% .github/workflows/oracle-octave.yml.  Fixtures: tests/golden/octave/*.json (scripts/export_fixtures_json.py; arrays
% flattened column-major next to their MATLAB shape).  Tolerance: the .m files are written for clarity, not in the C
% restatement's operation order, so iterates are held to 1e-9 relative (north_star) and iteration counts to equality.
%
% usage:  octave --no-gui --eval "cd oracle; check_octave"

function check_octave()
    here = fileparts(mfilename('fullpath'));
    gold = fullfile(here, '..', 'tests', 'golden', 'octave');
    nfail = 0;

    % ---- config 1 through admm_ocp.m
    d = jsondecode(fileread(fullfile(gold, 'cfg1_single_n20.json')));
    prob = struct('A', arr(d.prob.A), 'B', arr(d.prob.B), 's0', arr(d.prob.s0), ...
                  'block_type', int32(d.prob.block_type(:)), 'block_par', arr(d.prob.block_par));
    [x, z, u, hist] = admm_ocp(prob, d.opts);
    nfail = nfail + report('cfg1 iters', double(hist.iters(:)), double(d.out.iters(:)), 0);
    nfail = nfail + report('cfg1 x', x, arr(d.out.x), 1e-9);
    nfail = nfail + report('cfg1 z', z, arr(d.out.z), 1e-9);
    nfail = nfail + report('cfg1 u', u, arr(d.out.u), 1e-9);

    % ---- SCP fixture through admm_scp.m
    d = jsondecode(fileread(fullfile(gold, 'scp_b6_n12.json')));
    prob = struct('N', d.prob.N, 's0', arr(d.prob.s0), 'block_type', int32(d.prob.block_type(:)), ...
                  'block_par', arr(d.prob.block_par));
    [x, z, u, hist] = admm_scp(prob, d.opts, d.scp);
    nfail = nfail + report('scp passes', double(hist.scp_passes(:)), double(d.out.passes(:)), 0);
    nfail = nfail + report('scp iters_total', double(hist.scp_iters_total(:)), double(d.out.iters_total(:)), 0);
    nfail = nfail + report('scp x', x, arr(d.out.x), 1e-9);
    nfail = nfail + report('scp z', z, arr(d.out.z), 1e-9);

    if nfail > 0, error('check_octave: %d comparison(s) failed', nfail); end
    disp('check_octave: all comparisons passed');
end

function a = arr(s)
    shp = s.shape(:).';
    if numel(shp) == 1, shp = [shp 1]; end
    a = reshape(s.data(:), shp);
end

function bad = report(name, got, want, tol)
    if tol == 0
        bad = ~isequal(got(:), want(:));
        fprintf('%-18s %s\n', name, ternary(bad, 'DIFFERS', 'equal'));
    else
        err = max(abs(got(:) - want(:))) / max(1, max(abs(want(:))));
        bad = ~(err <= tol);
        fprintf('%-18s max rel diff %.3e (tol %.0e) %s\n', name, err, tol, ternary(bad, 'FAIL', 'ok'));
    end
    bad = double(bad);
end

function s = ternary(c, a, b)
    if c, s = a; else, s = b; end
end
