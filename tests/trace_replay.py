"""Replays the committed reference call traces (tests/golden/ref_trace_*.json, produced by
tests/golden/make_traces.py from the reference's own kzg.py / fft_ff.py) through any object that
has the reference's surface: commit(ck, polys), open(ck, polys, z, xi), fft_ff, ifft_ff,
fft_ff_interpolation.  Used by the CPU suite with the oracle and by the GPU suite with the
drop-in modules, so both are held to the same recorded reference outputs."""
import json
import os

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def H(v):
    return int(v, 16)


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def replay(trace, *, make_key, commit, open_, fft, ifft, interp, to_affine, to_ints):
    """Returns the number of records checked.  `make_key(affine_points)` builds a ck for the
    implementation under test; `to_affine(point)` -> (x, y) ints or None; `to_ints(seq)` -> ints."""
    keys = [make_key([None if p is None else (H(p[0]), H(p[1])) for p in k]) for k in trace.get("keys", [])]
    done = 0
    for rec in trace["calls"]:
        fn = rec["fn"]
        if fn == "commit":
            polys = [[H(c) for c in p] for p in rec["polys"]]
            got = [to_affine(c) for c in commit(keys[rec["ck_id"]], polys)]
            exp = [None if o is None else (H(o[0]), H(o[1])) for o in rec["out"]]
            assert got == exp, f"commit record {done}"
        elif fn == "open":
            polys = [[H(c) for c in p] for p in rec["polys"]]
            got = to_affine(open_(keys[rec["ck_id"]], polys, H(rec["z"]), H(rec["xi"])))
            exp = None if rec["out"] is None else (H(rec["out"][0]), H(rec["out"][1]))
            assert got == exp, f"open record {done}"
        else:
            vals = [H(v) for v in rec["in"]]
            f = {"fft_ff": fft, "ifft_ff": ifft, "fft_ff_interpolation": interp}[fn]
            got = to_ints(f(vals, H(rec["w"])))
            exp = [H(v) for v in rec["out"]]
            if fn == "fft_ff_interpolation":                 # a polynomial: trailing zeros stripped
                while got and got[-1] == 0:
                    got.pop()
            assert got == exp, f"{fn} record {done} (n={rec['n']})"
        done += 1
    return done
