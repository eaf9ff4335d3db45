"""Generates egt_b200/csrc/wide_bwd_program.cuh: the whole tcgen05.mma program of one backward handshake (the products
that consume key j's operands + the inputs of the group's next key + the commit) as ONE inline-asm statement per
kernel instantiation and compute group.

Why generated asm: the issuing warp, not the tensor core, paces the wide backward (cost model).  Here one
elected lane runs the program (CUTLASS' elect_one_sync pattern), every tensor-memory address is a literal (the group
index is a template parameter), descriptors are 64-bit registers advanced by immediates, and the k-step guards of the
row-contracted chains are forward branches: ~5 SASS instructions per tcgen05.mma instead of ~13 with one statement per
k-chain.

    python tools/gen/gen_wide_bwd_program.py > egt_b200/csrc/wide_bwd_program.cuh
"""

HI_SW = (1024 >> 4) | (1 << 14) | (2 << 29)          # desc_hi(1024, LAYOUT_SW128)
HI_NONE = (128 >> 4) | (1 << 14)                     # desc_hi(128, LAYOUT_NONE)
HI_TIMG = (2048 >> 4) | (1 << 14)                    # desc_hi(2048, LAYOUT_NONE)


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


class Prog:
    def __init__(self):
        self.lines, self.ops = [], []

    def op(self, name):
        if name not in self.ops:
            self.ops.append(name)
        return '%%%d' % self.ops.index(name)

    def emit(self, s):
        self.lines.append(s)


def gen(name, q, H, DK, DE, znone):
    D = H * DK
    DKS = D // 16
    EGN, DEP, DEW = 2 * H, max(DE, 16), max(DE, 16)
    EK = DEW // 16
    WN = DEP * (2 if DE < 16 else 1)
    TM_DQ, TM_W1, TM_W2, TM_G = 0, D, D + WN, D + 2 * WN
    GC = 80 + EGN + DEP
    G_S, G_DA, G_EG, G_HX, G_DX, G_T = 0, 16, 32, 32 + EGN, 48 + EGN, 48 + EGN + DEP
    tg = TM_G + q * GC                                  # literal: a 512-column allocation starts at column 0
    p = Prog()
    o, e = p.op, p.emit
    wcol = o('wcol') if DE < 16 else None                # d_e = 8: even / odd keys accumulate apart
    loKc_mn = o('loKc_mn'); ldc = o('ldc'); loWdx = o('loWdx'); loI = o('loI')
    loQm = o('loQm'); loDOm = o('loDOm'); loS = o('loS'); loA = o('loA'); loZ = o('loZ'); we = o('we'); wd = o('wd')
    first = o('first'); first_w = o('first_w'); KR = o('KR'); has_next = o('has_next')
    loQ = o('loQ'); loDO = o('loDO'); loKn = o('loKn'); loVn = o('loVn'); len_ = o('len'); ldn = o('ldn')
    loWeg = o('loWeg'); loWhx = o('loWhx'); use_lo = o('use_lo'); bar = o('bar')
    e('.reg .pred pe, pn, pt, pz, pacc, paccw, pk, plo;')
    e('.reg .b32 hsw, hno, hti, hz, iN16, iEG, iDX, iDQ, iT, iW, rd, ra;')
    e('.reg .b64 da, db;')
    e('elect.sync _|pe, 0xffffffff;')
    e('@!pe bra LDONE;')                                 # one elected lane runs the program
    e(f'setp.ne.b32 pn, {has_next}, 0;')
    e(f'setp.eq.b32 pt, {KR}, {KR};')
    e(f'setp.ne.b32 pz, {KR}, {KR};')
    e(f'setp.eq.b32 pacc, {first}, 0;')
    e(f'setp.eq.b32 paccw, {first_w}, 0;')
    e(f'setp.eq.b32 plo, {use_lo}, 0;')
    for reg, val in (('hsw', HI_SW), ('hno', HI_NONE), ('hti', HI_TIMG), ('hz', HI_TIMG if znone else HI_SW),
                     ('iN16', idesc(128, 16, 0, 0)), ('iEG', idesc(128, EGN, 0, 0)), ('iDX', idesc(128, DEP, 0, 0)),
                     ('iDQ', idesc(128, D, 0, 1)), ('iT', idesc(128, 16, 1, 1)), ('iW', idesc(128, DEP, 1, 1))):
        e(f'mov.b32 {reg}, {val};')

    def ss(d, acc, idr):
        e(f'tcgen05.mma.cta_group::1.kind::f16 [{d}], da, db, {idr}, {acc};')

    def ts(d, a, acc, idr):
        e(f'tcgen05.mma.cta_group::1.kind::f16 [{d}], [{a}], db, {idr}, {acc};')

    # ---- products of key j ----
    e(f'mov.b32 rd, {TM_DQ};')                           # dQ += dS Kexp
    e(f'mov.b32 ra, {tg + G_DA};')
    e(f'mov.b64 db, {{{loKc_mn}, hsw}};')
    ts('rd', 'ra', 'pacc', 'iDQ')
    e(f'mov.b64 da, {{{ldc}, hsw}};')                    # T = de' I + (r dZ) W'^T
    e(f'mov.b64 db, {{{loI}, hno}};')
    for s in range(DEP // 16):
        e(f'mov.b32 rd, {tg + G_DX + 16 * s};')
        if s and DE >= 16:
            e('add.u64 da, da, 2;')
        ss('rd', 'pz', 'iN16')
    e(f'mov.b32 rd, {tg + G_DX};')
    e(f'mov.b64 db, {{{loWdx}, hno}};')
    for s in range(EGN // 16):
        e(f'mov.b32 ra, {tg + G_S + 8 * s};')
        if s:
            e(f'add.u64 db, db, {2 * DEP};')
        ts('rd', 'ra', 'pt', 'iDX')
    # dK^T, dV^T: contraction over the query rows (k-steps >= KR skipped)
    for which, (loX, img, col) in enumerate(((loQm, loS, G_T), (loDOm, loA, G_T + 16))):
        e(f'mov.b32 rd, {tg + col};')
        e(f'mov.b64 da, {{{loX}, hsw}};')
        e(f'mov.b64 db, {{{img}, hti}};')
        ss('rd', 'pz', 'iT')
        for s in range(1, 8):
            e(f'setp.le.u32 pk, {KR}, {s};')
            e(f'@pk bra LT{which};')
            e('add.u64 da, da, 128;')
            e('add.u64 db, db, 16;')
            ss('rd', 'pt', 'iT')
        e(f'LT{which}:')
    # weight-gradient accumulators
    zstep = 16 if znone else 128
    for wi, (base, win) in enumerate(((TM_W1, we), (TM_W2, wd))):
        if wcol:
            e(f'add.u32 rd, {wcol}, {base};')
        else:
            e(f'mov.b32 rd, {base};')
        e(f'mov.b64 da, {{{loZ}, hz}};')
        e(f'mov.b64 db, {{{win}, hsw}};')
        ss('rd', 'paccw', 'iW')
        for s in range(1, 8):
            e(f'setp.le.u32 pk, {KR}, {s};')
            e(f'@pk bra LW{wi};')
            e(f'add.u64 da, da, {zstep};')
            e('add.u64 db, db, 128;')
            ss('rd', 'pt', 'iW')
        e(f'LW{wi}:')
    # ---- inputs of the group's next key ----
    e('@!pn bra LCOMMIT;')
    for col, loX, loB in ((G_S, loQ, loKn), (G_DA, loDO, loVn)):
        e(f'mov.b32 rd, {tg + col};')
        e(f'mov.b64 da, {{{loX}, hsw}};')
        e(f'mov.b64 db, {{{loB}, hsw}};')
        for s in range(DKS):
            if s:
                e(f'add.u64 da, da, {2 if s & 3 else 1024 - 6};')
                e(f'add.u64 db, db, {2 if s & 3 else 128 - 6};')
            ss('rd', 'pz' if s == 0 else 'pt', 'iN16')
    e(f'mov.b32 rd, {tg + G_EG};')
    e(f'mov.b64 db, {{{loWeg}, hno}};')
    for s in range(2 * EK):                               # W' = hi (+ lo when the logits are large)
        if s == EK:
            e('@plo bra LEGDONE;')
        if s % EK == 0:
            e(f'mov.b64 da, {{{len_}, hsw}};')
        else:
            e('add.u64 da, da, 2;')
        if s:
            e(f'add.u64 db, db, {2 * EGN};')
        ss('rd', 'pz' if s == 0 else 'pt', 'iEG')
    e('LEGDONE:')
    e(f'mov.b32 rd, {tg + G_HX};')
    e(f'mov.b64 da, {{{ldn}, hsw}};')
    e(f'mov.b64 db, {{{loWhx}, hno}};')
    for s in range(EK):
        if s:
            e('add.u64 da, da, 2;')
            e('add.u64 db, db, 32;')
        ss('rd', 'pz' if s == 0 else 'pt', 'iN16')
    e('LCOMMIT:')
    e(f'tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [{bar}];')
    e('LDONE:')
    n_mma = sum(1 for l in p.lines if 'tcgen05.mma' in l)
    args = ', '.join(f'uint32_t {n}' for n in p.ops)
    body = '\n'.join(f'      "{l}\\n\\t"' for l in p.lines)
    cons = ', '.join(f'"r"({n})' for n in p.ops)
    return f'''// {name}, compute group {q}: h = {H}, dk = {DK}, d_e = {DE}: {n_mma} tcgen05.mma + 1 commit
__device__ __forceinline__ void wide_bwd_program_{name}_g{q}({args}) {{
  asm volatile(
      "{{\\n\\t"
{body}
      "}}"
      ::{cons}
      : "memory");
}}
'''


print('// wide_bwd_program.cuh -- GENERATED by tools/gen/gen_wide_bwd_program.py; do not edit.')
print('//')
print('// The tcgen05.mma program of one backward handshake of wide_bwd.cu as ONE inline-asm statement per instantiation and')
print('// compute group (operands: warp-uniform 32-bit values; every lane of the converged issuer warp executes the statement,')
print('// one elected lane runs the program).  Tensor-memory addresses are literals: the kernel checks that its 512-column')
print('// allocation starts at column 0.')
print('#pragma once')
print('#include <stdint.h>')
print()
print('namespace egt {')
print()
for q in (0, 1):
    print(gen('c5', q, 16, 8, 32, False))
    print(gen('c3', q, 8, 12, 8, False))
    print(gen('c1', q, 8, 8, 64, True))
print('}  // namespace egt')
