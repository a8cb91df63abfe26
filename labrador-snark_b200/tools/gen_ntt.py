#!/usr/bin/env python3
"""Generates csrc/lab_ntt_gen.cuh: twiddle tables and straight-line, fully unrolled 32-point
transforms over F_{Q^2} (Q = 2^13 - 1) with *bound-tracked lazy reduction*.

R_q = F_Q[X]/(X^64+1) is isomorphic to F_{Q^2}[X]/(X^32 - i) via g_d = f_d + i f_{d+32}
(F_{Q^2} = F_Q[i], i^2 = -1, because Q = 3 mod 4), and X^32 - i splits completely over F_{Q^2}
because 128 | Q^2 - 1.  zeta = 2620 + 936 i is a primitive 128th root of unity with zeta^32 = i.
The forward transform is a merged-twiddle Cooley-Tukey network (node k has modulus
X^(2 len) - zeta^e, twiddle zeta^(e/2), children e/2 and e/2 + 64); slot j ends up holding
f(zeta^{e_j}), e_j = 1 mod 4.  The inverse is the Gentleman-Sande mirror with 32^-1 = 2^8.

Every value is a u32 holding a not-necessarily-canonical residue; the generator tracks an upper
bound per value and emits a fold (x & Q) + (x >> 13) only where a product sum could exceed 2^32,
so the emitted code is overflow-free by construction (asserted here, and checked bit-exactly
against the oracle in tests/).
"""
import os
import sys

Q = 8191
ZETA = (2620, 936)


def cmul(x, y):
    return ((x[0] * y[0] - x[1] * y[1]) % Q, (x[0] * y[1] + x[1] * y[0]) % Q)


def cpow(x, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = cmul(r, x)
        x = cmul(x, x)
        e >>= 1
    return r


def tables():
    node_e = {1: 32}
    tw_exp = {}
    for k in range(1, 32):
        tw_exp[k] = node_e[k] // 2
        node_e[2 * k] = node_e[k] // 2
        node_e[2 * k + 1] = node_e[k] // 2 + 64
    fwd = {k: cpow(ZETA, tw_exp[k]) for k in range(1, 32)}
    inv = {k: cpow(ZETA, 128 - tw_exp[k]) for k in range(1, 32)}
    slot = [node_e[32 + j] for j in range(32)]
    return fwd, inv, slot


class Emitter:
    def __init__(self):
        self.lines = []
        self.bound = {}
        self.nfold = 0
        self.nmul = 0

    def emit(self, s):
        self.lines.append("    " + s)

    def fold(self, v):
        b = self.bound[v]
        self.emit(f"{v} = lab_fold({v});")
        self.bound[v] = Q + (b >> 13) if b > Q else b       # (x & Q) <= Q, x >> 13 <= b >> 13
        self.nfold += 1

    def fold_to(self, v, limit):
        while self.bound[v] > limit:
            before = self.bound[v]
            self.fold(v)
            assert self.bound[v] < before, "fold made no progress"

    def canon(self, v):
        # -> [0, Q)
        self.fold_to(v, 2 * Q - 1)
        self.emit(f"{v} = lab_csub({v});")
        self.bound[v] = Q - 1


def cmul_const(E, dst_re, dst_im, a_re, a_im, c, tfolds=2):
    """(dst) = (a) * c, c = (cr, ci) canonical constants; result folded `tfolds` times (2 => < 2*Q)."""
    cr, ci = c
    nci = (Q - ci) % Q
    # make sure the product sums fit in 32 bits
    limit = (1 << 32) - 1
    while True:
        worst_re = E.bound[a_re] * cr + E.bound[a_im] * nci
        worst_im = E.bound[a_re] * ci + E.bound[a_im] * cr
        if max(worst_re, worst_im) <= limit:
            break
        v = a_re if E.bound[a_re] >= E.bound[a_im] else a_im
        E.fold(v)
    E.emit(f"{dst_re} = {a_re} * {cr}u + {a_im} * {nci}u;")
    E.emit(f"{dst_im} = {a_re} * {ci}u + {a_im} * {cr}u;")
    E.nmul += 4
    E.bound[dst_re] = worst_re
    E.bound[dst_im] = worst_im
    if tfolds >= 2:
        E.fold_to(dst_re, 2 * Q - 1)
        E.fold_to(dst_im, 2 * Q - 1)
    else:
        E.fold(dst_re)
        E.fold(dst_im)


def kmult(bound):
    """smallest multiple of Q that is >= bound"""
    return ((bound + Q - 1) // Q) * Q


IN_BOUND = (1 << 14) - 1      # callers guarantee 14-bit inputs (canonical input passes as is; anything larger is folded twice first)


def finish(E, lazy):
    for j in range(32):
        for v in (f"re[{j}]", f"im[{j}]"):
            if lazy:
                E.fold_to(v, 2 * Q - 1)      # < 2Q: good enough for lab_cmul operands and for packing into 16 bits
            else:
                E.canon(v)


def gen_fwd(fwd, lazy=False, policy=(2, 2, 2, 2, 2)):
    E = Emitter()
    for j in range(32):
        E.bound[f"re[{j}]"] = IN_BOUND
        E.bound[f"im[{j}]"] = IN_BOUND
    E.emit("uint32_t tr, ti;")
    k = 1
    length = 16
    stage = 0
    while length >= 1:
        for start in range(0, 32, 2 * length):
            c = fwd[k]
            k += 1
            for j in range(start, start + length):
                lo_r, lo_i, hi_r, hi_i = f"re[{j}]", f"im[{j}]", f"re[{j+length}]", f"im[{j+length}]"
                cmul_const(E, "tr", "ti", hi_r, hi_i, c, policy[stage])
                for lo, hi, t in ((lo_r, hi_r, "tr"), (lo_i, hi_i, "ti")):
                    K = kmult(E.bound[t])
                    bl = E.bound[lo]
                    assert bl + K < (1 << 32)
                    E.emit(f"{hi} = {lo} + {K}u - {t};")
                    E.emit(f"{lo} = {lo} + {t};")
                    E.bound[hi] = bl + K
                    E.bound[lo] = bl + E.bound[t]
        length //= 2
        stage += 1
    finish(E, lazy)
    return E


def gen_inv(inv, policy=(2, 2, 2, 2, 2)):
    E = Emitter()
    for j in range(32):
        E.bound[f"re[{j}]"] = IN_BOUND
        E.bound[f"im[{j}]"] = IN_BOUND
    E.emit("uint32_t ur, ui;")
    length = 1
    stage = 0
    while length <= 16:
        k = 16 // length
        for start in range(0, 32, 2 * length):
            c = inv[k]
            k += 1
            if length == 16:
                c = cmul(c, (256, 0))            # fold 32^-1 = 2^8 into the last level
            for j in range(start, start + length):
                lo_r, lo_i, hi_r, hi_i = f"re[{j}]", f"im[{j}]", f"re[{j+length}]", f"im[{j+length}]"
                for lo, hi, u in ((lo_r, hi_r, "ur"), (lo_i, hi_i, "ui")):
                    # keep the difference operand small enough for the following multiplication
                    E.fold_to(hi, 1 << 17)
                    E.fold_to(lo, 1 << 17)
                    K = kmult(E.bound[hi])
                    E.emit(f"{u} = {lo} + {K}u - {hi};")
                    E.bound[u] = E.bound[lo] + K
                    E.emit(f"{lo} = {lo} + {hi};")
                    E.bound[lo] = E.bound[lo] + E.bound[hi]
                cmul_const(E, hi_r, hi_i, "ur", "ui", c, policy[stage])
        length *= 2
        stage += 1
    for j in range(16):                        # the sum halves still need the 2^8 factor
        for v in (f"re[{j}]", f"im[{j}]"):
            E.fold_to(v, (1 << 24) - 1)
            E.emit(f"{v} = {v} << 8;")
            E.bound[v] = E.bound[v] << 8
    for j in range(32):
        E.canon(f"re[{j}]")
        E.canon(f"im[{j}]")
    return E


def main():
    fwd, inv, slot = tables()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "lab_ntt_gen.cuh")
    if len(sys.argv) > 1:
        out = sys.argv[1]
    import itertools

    def best(fn, **kw):
        cands = []
        for pol in itertools.product((1, 2), repeat=5):
            try:
                e = fn(policy=pol, **kw)
            except AssertionError:
                continue
            cands.append((e.nfold, pol, e))
        cands.sort(key=lambda t: (t[0], t[1]))
        return cands[0][2], cands[0][1]

    (F, pf), (FL, pfl), (I, pi_) = best(lambda policy: gen_fwd(fwd, policy=policy)), best(lambda policy: gen_fwd(fwd, lazy=True, policy=policy)), best(lambda policy: gen_inv(inv, policy=policy))
    print("fold policies (folds of the twiddle product per stage):", pf, pfl, pi_)
    pk = lambda c: c[0] | (c[1] << 16)
    with open(out, "w") as f:
        f.write("// GENERATED by tools/gen_ntt.py -- do not edit.  See that file for the math.\n")
        f.write("#pragma once\n#include <stdint.h>\n\n")
        f.write(f"// forward: {F.nmul} multiplies, {F.nfold} folds; inverse: {I.nmul} multiplies, {I.nfold} folds\n")
        f.write("// twiddle of tree node k (1..31), packed re | im << 16; index 0 unused\n")
        f.write("#define LAB_TW_FWD_INIT {0u, " + ", ".join(f"{pk(fwd[k])}u" for k in range(1, 32)) + "}\n")
        f.write("#define LAB_TW_INV_INIT {0u, " + ", ".join(f"{pk(inv[k])}u" for k in range(1, 32)) + "}\n")
        f.write("// slot j holds f(zeta^e_j)\n")
        f.write("#define LAB_SLOT_EXP_INIT {" + ", ".join(str(e) for e in slot) + "}\n\n")
        f.write(f"// in: residues <= {IN_BOUND} (14 bits); out: canonical [0,Q)\n")
        f.write("__device__ __forceinline__ void lab_ntt32_fwd_regs(uint32_t (&re)[32], uint32_t (&im)[32]) {\n")
        f.write("\n".join(F.lines))
        f.write("\n}\n\n")
        f.write("// same, but outputs are only reduced to < 2Q (operands of lab_cmul / 16-bit packing)\n")
        f.write("__device__ __forceinline__ void lab_ntt32_fwd_regs_lazy(uint32_t (&re)[32], uint32_t (&im)[32]) {\n")
        f.write("\n".join(FL.lines))
        f.write("\n}\n\n")
        f.write(f"// in: residues <= {IN_BOUND} in slot order; out: canonical coefficients g_d = f_d + i f_{{d+32}}, scaled by 1/32\n")
        f.write("__device__ __forceinline__ void lab_ntt32_inv_regs(uint32_t (&re)[32], uint32_t (&im)[32]) {\n")
        f.write("\n".join(I.lines))
        f.write("\n}\n")
    print(f"wrote {out}: fwd {F.nmul} mul {F.nfold} folds, inv {I.nmul} mul {I.nfold} folds")


if __name__ == "__main__":
    main()
