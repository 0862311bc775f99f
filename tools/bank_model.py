import sys
def sim(N, TL, rad, esz, L, cols, padfn, NPAD):
    NP = len(rad); T = TL * L
    def slot(pos, l): return padfn(pos) * L + l if cols else l * NPAD + padfn(pos)
    def il(tid): return (tid // L, tid % L) if cols else (tid % TL, tid // TL)
    def wf(instrs):
        tot = ideal = 0
        grp = 128 // esz
        for instr in instrs:
            for g0 in range(0, len(instr), grp):
                g = [a for a in instr[g0:g0+grp] if a is not None]
                if not g: continue
                banks = {}
                for a in g:
                    w0 = a * esz // 4
                    for wd in range(esz // 4):
                        banks.setdefault((w0 + wd) % 32, set()).add((w0 + wd) // 32)
                tot += max(len(v) for v in banks.values()); ideal += 1
        return tot, ideal
    tot = idl = 0; before = 1
    for p in range(NP):
        r = rad[p]; NB = N // r; G = -(-NB // TL); P = before
        for kind in (["w"] if p < NP - 1 else []) + (["r"] if p > 0 else []):
            instrs = []
            for m in range(G):
                for q in range(r):
                    row = []
                    for tid in range(T):
                        i, l = il(tid); b = i + TL * m
                        if b >= NB: row.append(None); continue
                        k = b % P
                        row.append(slot((b - k) * r + k + q * P, l) if kind == "w" else slot(b + q * NB, l))
                    instrs.append(row)
            t, i_ = wf(instrs); tot += t; idl += i_
        before *= r
    return tot, idl
N, TL, rad, esz, L, cols = int(sys.argv[1]), int(sys.argv[2]), [int(x) for x in sys.argv[3].split(".")], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
R0 = rad[0]
out = []
for name, D in (("R0", R0), ("none", 0), ("8", 8), ("16", 16), ("32", 32), ("R0*2", 2 * R0)):
    padfn = (lambda a, D=D: a + a // D) if D else (lambda a: a)
    NPAD = N + (N // D if D else 0)
    t, i = sim(N, TL, rad, esz, L, cols, padfn, NPAD)
    out.append("%s x%.2f" % (name, t / i))
print(" ".join(sys.argv[1:]), "|", "  ".join(out))
