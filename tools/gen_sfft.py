#!/usr/bin/env python3
"""Generates the instantiation lists of the fast Stockham kernels (ndrustfft_b200/csrc/{sfft,rsfft}_inst_*.cu).

Run `python tools/gen_sfft.py` after changing the schedule rules; the generated files are committed.
One entry = (dtype, N, threads-per-lane TL, radices, lanes per CTA L, rows|cols layout, min CTAs/SM, family).

Two schedule families are instantiated and the host picks per case (csrc/ndfft_b200.cu: find_sfft / find_rsfft):
  family A: few passes, many registers  (radix 16, f32 last pass radix 32; E = 16/32 points per thread)
  family B: more CTAs per SM            (radix 8 for f64, 16 for f32;      E = 8/16 points per thread)
Measured on B200 (profiles/r1f_tune*.jsonl): B wins wherever occupancy is the limiter (e.g. 512-point f64 rows
81 % vs 64 % of the HBM roofline).

Also generated: the fused Bluestein kernels (bsfft_inst_*.cu), the fused two-pass column kernels (fs2_inst.cu, opt-in at
run time) and, for the mixed-radix lengths, register-capped variants of the same tiles (capped_variants: twice the CTAs
per SM for 8-16 bytes of spill, 1.2-1.5x faster; the host picks them through its resident-thread rule).
`NDFB_GEN_EXPERIMENT=1` adds extra-CTA variants of the power-of-two kernels for A/B runs (they lose; not committed).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "ndrustfft_b200", "csrc")

MIXED = {  # N: (TL, radices)   ({2,3,5}-smooth lengths of BASELINE config c5a and friends)
    360: (60, [6, 6, 10]),
    1000: (100, [10, 10, 10]),
    384: (64, [6, 8, 8]),
    100: (10, [10, 10]),
    36: (6, [6, 6]),
    60: (10, [6, 10]),
    216: (36, [6, 6, 6]),
    600: (100, [6, 10, 10]),
    264: (33, [8, 3, 11]),        # benches/ndrustfft.rs FFT_SIZES / DCT_SIZES (264, 265): 2^3 3 11
    132: (33, [4, 3, 11]),        # core of the 264-point r2c
    81: (9, [9, 9]),
    729: (81, [9, 9, 9]),
    4095: (512, [13, 9, 7, 5]),   # DCT-I of 4096 points (BASELINE c4): core 4095 = 3^2 5 7 13
}


def pow2_schedule(N, f64, fam):
    if fam == 0:
        emax, rbase = (16 if f64 else 32), 16
    else:
        emax, rbase = (8 if f64 else 16), (8 if f64 else 16)
    if N <= emax:
        return 1, [N]
    rad, rem = [], N
    while rem > 1:
        r = min(rbase, rem)
        rad.append(r)
        rem //= r
    if fam == 0 and not f64 and len(rad) >= 2 and rad[-1] == 2 and rad[-2] == 16:
        rad = rad[:-2] + [32]            # f32: one radix-32 last pass instead of 16 then 2
    if len(rad) > 4:
        return None
    return N // max(rad), rad


def schedule(N, f64, fam):
    if N in MIXED:
        return MIXED[N] if fam == 0 else None
    return pow2_schedule(N, f64, fam)


def npad(N, r0):
    return N if r0 % 2 else N + N // r0       # Sched::NPAD: an odd first radix needs no pad


def make(f64, N, TL, rad, L, cols, fam, minb=None, always_smem=False):
    cs = 16 if f64 else 8
    T = TL * L
    smem = L * npad(N, rad[0]) * cs if (len(rad) > 1 or always_smem) else 0
    E = max(-(-(N // r) // TL) * r for r in rad)
    if minb is None:
        regs = E * (4 if f64 else 2) + (32 if fam >= 1 else (56 if f64 else 44))
        regs = min(regs, 255)
        minb = max(1, min(65536 // (T * regs), (227 * 1024) // max(smem, 1) if smem else 8, 8))
    return dict(f64=f64, N=N, TL=TL, rad=rad + [1] * (4 - len(rad)), L=L, cols=cols, minb=minb, T=T, smem=smem, E=E, fam=fam)


def rows_L(N, TL, r0, cs):
    L = max(1, 256 // TL)
    while L > 1 and L * npad(N, r0) * cs > 72 * 1024:
        L //= 2
    p = 1
    while p * 2 <= L:
        p *= 2
    return p


def capped_variants(entries, f64, allow_4095=False):
    """Mixed-radix butterflies take ~120 (f64) / ~80 (f32) registers when allowed to; a variant of the same tile capped
    for twice the CTAs per SM spills 8-16 bytes and runs 1.2-1.5x faster (profiles/r1z_tune_mixed.jsonl): 360-point c128
    columns 48 % -> 74 % of the roofline.  The host prefers it through the resident-thread rule."""
    out = []
    for e in entries:
        regs_cap = 65536 // (e["T"] * e["minb"] * 2)
        if (allow_4095 or e["N"] != 4095) and regs_cap >= (64 if f64 else 48) and e["T"] * e["minb"] * 2 <= 2048:
            out.append(dict(e, minb=e["minb"] * 2))
    return out


def c2c_rows(f64):
    out = []
    cs = 16 if f64 else 8
    for fam in (0, 1):
        for N in [64, 128, 256, 512, 1024, 2048, 4096, 8192] + sorted(MIXED):
            sc = schedule(N, f64, fam)
            if sc is None:
                continue
            TL, rad = sc
            if TL > 512 or TL < 2:
                continue
            e = make(f64, N, TL, rad, rows_L(N, TL, rad[0], cs), 0, fam)
            if e["smem"] <= 220 * 1024 and 32 <= e["T"] <= 1024:
                out.append(e)
                if N in MIXED:
                    out.extend(capped_variants([e], f64))
                elif fam == 1 and N >= 2048 and os.environ.get("NDFB_GEN_EXPERIMENT"):
                    # one more CTA per SM if shared memory allows it (experiment: NDFB_SFFT_PICK selects it)
                    mb = e["minb"] + 1
                    if e["smem"] * mb <= 227 * 1024 and e["T"] * mb <= 2048:
                        out.append(dict(e, minb=mb))
    return dedup(out)


def c2c_cols(f64):
    out = []
    cs = 16 if f64 else 8
    for fam in (0, 1):
        for N in [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192] + sorted(MIXED):
            sc = schedule(N, f64, fam)
            if sc is None:
                continue
            TL, rad = sc
            cand = []
            for L in (32, 16, 8, 4, 2):
                if L * cs > (256 if N <= 256 else 128):   # short columns: 256-byte rows keep enough loads in flight per CTA
                    continue
                e = make(f64, N, TL, rad, L, 1, fam)
                # family B keeps <= 64 registers, so a 1024-thread CTA (32 warps on a tile that owns the whole SM) is allowed
                tmax = 1024 if (fam == 1 and e["E"] * (4 if f64 else 2) <= 32) else 512
                if e["T"] > tmax or e["T"] < 32 or e["smem"] > 200 * 1024:
                    continue
                cand.append(e)
            out.extend(cand[:3])  # the three widest tiles that fit
            if N in MIXED:
                out.extend(capped_variants(cand[:2], f64))
    return dedup(out)


# core lengths whose family-B schedule is also instantiated small-radix-FIRST (4.8.8.8): pass 0 then has two butterflies per
# thread and the C2R / DCT-III prologue runs mirror-paired from registers (sfft_kernel.cuh: kMirrorPro)
REVERSED = {1: (256, 2048), 0: (2048,)}


def real_entries(f64):
    out = []
    rs = 8 if f64 else 4
    cs = 2 * rs
    for N in REVERSED[f64]:
        TL, rad = pow2_schedule(N, f64, 1)
        if rad[-1] >= rad[0]:
            continue
        rev = rad[::-1]
        out.append(make(f64, N, TL, rev, rows_L(N, TL, rev[0], cs), 0, 2, always_smem=True))
        nfit = 0
        for L in (32, 16, 8, 4, 2):
            if L * rs > 128:
                continue
            e = make(f64, N, TL, rev, L, 1, 2, always_smem=True)
            tmax = 1024 if e["E"] * (4 if f64 else 2) <= 32 else 512
            if e["T"] > tmax or e["T"] < 32 or e["smem"] > 200 * 1024:
                continue
            out.append(e)
            nfit += 1
            if nfit == 2:
                break
    for fam in (0, 1):
        for N in [64, 128, 256, 512, 1024, 2048, 4096, 4095, 132, 264]:
            sc = schedule(N, f64, fam)
            if sc is None:
                continue
            TL, rad = sc
            if TL < 2 or TL > 512:
                continue
            e = make(f64, N, TL, rad, rows_L(N, TL, rad[0], cs), 0, fam, always_smem=True)
            out.append(e)
            if N in MIXED:
                out.extend(capped_variants([e], f64, allow_4095=True))
            if fam == 1 and N >= 256 and N != 4095 and os.environ.get("NDFB_GEN_EXPERIMENT"):
                mb = e["minb"] + 1
                if e["smem"] * mb <= 227 * 1024 and e["T"] * mb <= 2048:
                    out.append(dict(e, minb=mb))
        for N in [32, 64, 128, 256, 512, 1024, 2048, 4096, 132, 264, 4095]:
            sc = schedule(N, f64, fam)
            if sc is None:
                continue
            TL, rad = sc
            nfit = 0
            for L in (32, 16, 8, 4, 2):
                if L * rs > 128:
                    continue
                e = make(f64, N, TL, rad, L, 1, fam, always_smem=True)
                tmax = 1024 if (fam == 1 and e["E"] * (4 if f64 else 2) <= 32) else 512
                if N == 4095:
                    # DCT-I of 4096 points along a strided axis (BASELINE c4): a two-column tile of 1024 threads capped at
                    # 64 registers beats the general kernel this length used before (8 % of the roofline)
                    tmax = 1024
                    e = dict(e, minb=1)
                if e["T"] > tmax or e["T"] < 32 or e["smem"] > 200 * 1024:
                    continue
                out.append(e)
                nfit += 1
                if nfit == 2:
                    break
    return dedup(out)


# alternative schedules of a core length, instantiated next to the MIXED one for the real kinds only (the host's entry rules or
# NDFB_RSFFT_PICK choose): 4095 = 15.13.7.3 on 320 threads wastes fewer thread slots per pass (85-98 % busy against 57-89 % for
# 13.9.7.5 on 512) and three CTAs fit an SM
ALT_REAL = {4095: (320, [15, 13, 7, 3])}


def alt_real_entries(f64):
    out = []
    cs = 16 if f64 else 8
    for N, (TL, rad) in ALT_REAL.items():
        for minb in (2, 3):
            out.append(make(f64, N, TL, rad, 1, 0, 0, minb=minb, always_smem=True))
        out.append(make(f64, N, TL, rad, 2, 1, 0, minb=1, always_smem=True))
    return out


def bluestein_entries(f64):
    """Fused Bluestein kernels: M-point family-B schedules (rows: one tile shape, columns: the two widest that fit)."""
    out = []
    cs = 16 if f64 else 8
    for M in [64, 128, 256, 512, 1024, 2048, 4096] + ([] if f64 else [8192]):
        sc = schedule(M, f64, 1)
        if sc is None:
            continue
        TL, rad = sc
        if TL < 2 or TL > 512:
            continue
        out.append(make(f64, M, TL, rad, rows_L(M, TL, rad[0], cs), 0, 1, always_smem=True))
        cand = []
        for L in (16, 8, 4, 2):
            if L * cs > 128:
                continue
            e = make(f64, M, TL, rad, L, 1, 1, always_smem=True)
            if e["T"] > 512 or e["T"] < 32 or e["smem"] > 200 * 1024:
                continue
            cand.append(e)
        out.extend(cand[:2])
    return dedup(out)


def fs2_entries():
    """Fused two-pass kernels (fs2_kernel): N = N1 * N2 strided columns, both passes family-B schedules run by 256-thread CTAs."""
    out = []
    for f64 in (0, 1):
        for N1, N2 in ((64, 128), (128, 128), (128, 256), (256, 256)):
            TL1, r1 = pow2_schedule(N1, f64, 1)
            TL2, r2 = pow2_schedule(N2, f64, 1)
            r1 = r1 + [1] * (4 - len(r1))
            r2 = r2 + [1] * (4 - len(r2))
            R = "double" if f64 else "float"
            for T in (256, 128, 512):     # first match is the default; NDFB_FS2_T picks another
                L1, L2 = T // TL1, T // TL2
                if min(L1, L2) * (16 if f64 else 8) < 64:
                    continue
                out.append(f"    FS2_ENTRY({R}, {f64}, {N1}, {TL1}, {r1[0]}, {r1[1]}, {r1[2]}, {r1[3]}, {L1}, "
                           f"{N2}, {TL2}, {r2[0]}, {r2[1]}, {r2[2]}, {r2[3]}, {L2}, {1024 // T}),")
    return out


def bulk_entries():
    """Persistent bulk-async row kernels (sfft_rows_bulk_kernel): family-B schedules for the long power-of-two rows."""
    out = []
    for f64, lengths in ((0, (1024, 2048, 4096, 8192)), (1, (512, 1024, 2048, 4096))):
        cs = 16 if f64 else 8
        R = "double" if f64 else "float"
        for N in lengths:
            TL, rad = pow2_schedule(N, f64, 1)
            rad = rad + [1] * (4 - len(rad))
            L = rows_L(N, TL, rad[0], cs)
            while L > 1 and L * (npad(N, rad[0]) + 2 * N) * cs + 16 > 110 * 1024:
                L //= 2
            smem = L * (npad(N, rad[0]) + 2 * N) * cs + 16
            T = TL * L
            minb = max(1, min((227 * 1024) // smem, 65536 // (T * 64), 4))
            out.append(f"    SFFT_BULK_ENTRY({R}, {f64}, {N}, {TL}, {rad[0]}, {rad[1]}, {rad[2]}, {rad[3]}, {L}, {minb}),  // T={T} smem={smem}")
    return out


def pipe_entries():
    """Software-pipelined persistent column kernels (pipe_kernel.cuh): tiles that own a whole SM (both passes of the 2^24-point
    rows of BASELINE c5b, the long f64 columns), plus a few small schedules for the emulator / GPU parity tests."""
    out = []
    prod = [(0, 4096, 4), (0, 8192, 2), (0, 2048, 8), (1, 2048, 4), (1, 4096, 2), (1, 1024, 8), (1, 1000, 4), (1, 1000, 8), (0, 4096, 2)]
    small = [(0, 64, 4), (0, 512, 4), (0, 256, 4), (1, 64, 2), (1, 512, 4), (1, 360, 2)]
    # one-lane tiles = contiguous rows in and out (opt-in, NDFB_PIPE=2): 64 KiB rows, two CTAs per SM
    rows = [(0, 8192, 1), (1, 4096, 1), (0, 4096, 1), (1, 2048, 1), (0, 256, 1), (1, 512, 1)]
    for f64, N, L in prod + small + rows:
        sc = schedule(N, f64, 1) or schedule(N, f64, 0)
        TL, rad = sc
        rad = list(rad) + [1] * (4 - len(rad))
        R = "double" if f64 else "float"
        cs = 16 if f64 else 8
        for inmode in ((1,) if L == 1 else (0, 1)):
            smem = L * (N + 4) * cs + L * npad(N, rad[0]) * cs // 2
            minb = max(1, min((227 * 1024) // smem, 65536 // (TL * L * 64), 4))
            out.append(f"    SFFT_PIPE_ENTRY({R}, {f64}, {N}, {TL}, {rad[0]}, {rad[1]}, {rad[2]}, {rad[3]}, {L}, {inmode}, {minb}),  // T={TL * L}")
    return out


def dedup(entries):
    seen, out = set(), []
    for e in entries:
        key = (e["f64"], e["N"], e["TL"], tuple(e["rad"]), e["L"], e["cols"], e["minb"])
        if key in seen:
            continue
        seen.add(key)
        out.append(e)
    return out


def fmt(e, macro, extra=""):
    R = "double" if e["f64"] else "float"
    r = e["rad"]
    return (f"    {macro}({R}, {e['f64']}, {extra}{e['N']}, {e['TL']}, {r[0]}, {r[1]}, {r[2]}, {r[3]}, {e['L']}, {e['cols']}, {e['minb']}, {e['fam']}),"
            f"  // T={e['T']} smem={e['smem']} E={e['E']}")


WRITTEN = set()


def write(path, lines):
    """Unchanged files keep their mtime (make would otherwise rebuild ~900 kernel instances)."""
    WRITTEN.add(path)
    full, text = os.path.join(OUT, path), "\n".join(lines) + "\n"
    if os.path.exists(full) and open(full).read() == text:
        return
    with open(full, "w") as f:
        f.write(text)


def main():
    total = 0
    head = ["// GENERATED by tools/gen_sfft.py — do not edit.", '#include "sfft_inst.h"', "", "namespace ndfb {", ""]
    groups = {"f32_rows": c2c_rows(0), "f32_cols": c2c_cols(0), "f64_rows": c2c_rows(1), "f64_cols": c2c_cols(1)}
    for name, ents in groups.items():
        write(f"sfft_inst_{name}.cu", head + [f"const SfftEntry kSfft_{name}[] = {{"] + [fmt(e, "SFFT_ENTRY") for e in ents] +
              ["};", f"const int kSfft_{name}_count = {len(ents)};", "", "}  // namespace ndfb"])
        total += len(ents)
    for f64 in (0, 1):
        nm = "f64" if f64 else "f32"
        ents = bluestein_entries(f64)
        write(f"bsfft_inst_{nm}.cu", head + [f"const BsfftEntry kBsfft_{nm}[] = {{"] + [fmt(e, "BSFFT_ENTRY") for e in ents] +
              ["};", f"const int kBsfft_{nm}_count = {len(ents)};", "", "}  // namespace ndfb"])
        total += len(ents)
    ents = bulk_entries()
    write("sfft_inst_bulk.cu", head + ["const SfftBulkEntry kSfftBulk[] = {"] + ents + ["};", f"const int kSfftBulk_count = {len(ents)};", "", "}  // namespace ndfb"])
    total += len(ents)
    ents = pipe_entries()
    write("sfft_inst_pipe.cu", ["// GENERATED by tools/gen_sfft.py — do not edit.", '#include "pipe_inst.h"', "", "namespace ndfb {", "",
                                "const SfftPipeEntry kSfftPipe[] = {"] + ents + ["};", f"const int kSfftPipe_count = {len(ents)};", "", "}  // namespace ndfb"])
    total += len(ents)
    ents = fs2_entries()
    write("fs2_inst.cu", head + ["const Fs2Entry kFs2[] = {"] + ents + ["};", f"const int kFs2_count = {len(ents)};", "", "}  // namespace ndfb"])
    total += len(ents)
    rnames = []
    for f64 in (0, 1):
        ents = real_entries(f64) + alt_real_entries(f64)
        for kind in ["RK_R2C", "RK_C2R", "RK_DCT1", "RK_DCT2", "RK_DCT3", "RK_DCT4"]:
            nm = f"{'f64' if f64 else 'f32'}_{kind[3:].lower()}"
            write(f"rsfft_inst_{nm}.cu", head + [f"const RsfftEntry kRsfft_{nm}[] = {{"] + [fmt(e, "RSFFT_ENTRY", kind + ", ") for e in ents] +
                  ["};", f"const int kRsfft_{nm}_count = {len(ents)};", "", "}  // namespace ndfb"])
            total += len(ents)
            rnames.append(nm)
    write("rsfft_tables.inc", ["// GENERATED by tools/gen_sfft.py — do not edit."] + [f"RSFFT_TABLE({nm})" for nm in rnames])
    for f in os.listdir(OUT):      # instance files of an earlier rule set
        if f.startswith(("sfft_inst_", "rsfft_inst_", "bsfft_inst_", "fs2_inst")) and f.endswith(".cu") and f not in WRITTEN:
            os.remove(os.path.join(OUT, f))
    print("generated", total, "instances")
    if "-v" in sys.argv:
        for name, ents in groups.items():
            for e in ents:
                print(name, e)


if __name__ == "__main__":
    main()
