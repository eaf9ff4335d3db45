"""Generates egt_b200/csrc/wide_bwd_program.cuh: the whole tcgen05.mma program of one backward handshake (the products
that consume key j's operands + the inputs of the group's next key + the commit) as ONE inline-asm statement per
kernel instantiation.  One election, operands moved to uniform registers once, descriptor stepping inside the asm:
~4 SASS instructions per tcgen05.mma instead of ~13 when every k-chain is its own statement (cost model).

    python tools/gen/gen_wide_bwd_program.py > egt_b200/csrc/wide_bwd_program.cuh
"""

HI_SW = (1024 >> 4) | (1 << 14) | (2 << 29)          # desc_hi(1024, LAYOUT_SW128)
HI_NONE = (128 >> 4) | (1 << 14)                     # desc_hi(128, LAYOUT_NONE)
HI_TIMG = (2048 >> 4) | (1 << 14)                    # desc_hi(2048, LAYOUT_NONE)


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


class Prog:
    def __init__(self):
        self.lines = []
        self.ops = []            # (name, c expression)

    def op(self, name):
        if name not in [n for n, _ in self.ops]:
            self.ops.append((name, name))
        return '%%%d' % [n for n, _ in self.ops].index(name)

    def emit(self, s):
        self.lines.append(s)

    def mma_ss(self, pred, d, alo, ahi, blo, bhi, idesc_reg, acc):
        self.emit(f'mov.b64 da, {{{alo}, {ahi}}};')
        self.emit(f'mov.b64 db, {{{blo}, {bhi}}};')
        self.emit(f'tcgen05.mma.cta_group::1.kind::f16 [{d}], da, db, {idesc_reg}, {acc};')

    def mma_ts(self, pred, d, a_tmem, blo, bhi, idesc_reg, acc):
        self.emit(f'mov.b64 db, {{{blo}, {bhi}}};')
        self.emit(f'tcgen05.mma.cta_group::1.kind::f16 [{d}], [{a_tmem}], db, {idesc_reg}, {acc};')


def gen(name, H, DK, DE, znone):
    D = H * DK
    DKS, NQA = D // 16, (D + 63) // 64
    EGN, DEP, DEW = 2 * H, max(DE, 16), max(DE, 16)
    EK = DEW // 16
    G_S, G_DA, G_EG, G_HX, G_DX, G_T = 0, 16, 32, 32 + EGN, 48 + EGN, 48 + EGN + DEP
    p = Prog()
    o = p.op
    # operands (all warp-uniform 32-bit values computed by the caller)
    tg = o('tg'); tm_dq = o('tm_dq'); tm_w1 = o('tm_w1'); tm_w2 = o('tm_w2')
    loKc_mn = o('loKc_mn'); ldc = o('ldc'); loWdx = o('loWdx'); loI = o('loI')
    loQm = o('loQm'); loDOm = o('loDOm'); loS = o('loS'); loA = o('loA'); loZ = o('loZ'); we = o('we'); wd = o('wd')
    first = o('first'); first_w = o('first_w'); KR = o('KR'); has_next = o('has_next')
    loQ = o('loQ'); loDO = o('loDO'); loKn = o('loKn'); loVn = o('loVn'); len_ = o('len'); ldn = o('ldn')
    loWeg = o('loWeg'); loWhx = o('loWhx'); bar = o('bar')
    e = p.emit
    e('.reg .pred pe, pn, pt, pz, pacc, paccw, pk;')
    e('.reg .b32 hsw, hno, hti, hz, iN16, iEG, iDX, iDQ, iT, iW, ra, rb, rd;')
    e('.reg .b64 da, db;')
    e('elect.sync _|pe, 0xffffffff;')
    e('@!pe bra LDONE;')                                   # one elected lane runs the program (CUTLASS' elect_one_sync pattern)
    e(f'setp.ne.b32 pn, {has_next}, 0;')
    e(f'setp.eq.b32 pt, {tg}, {tg};')
    e(f'setp.ne.b32 pz, {tg}, {tg};')
    e(f'setp.eq.b32 pacc, {first}, 0;')
    e(f'setp.eq.b32 paccw, {first_w}, 0;')
    e(f'mov.b32 hsw, {HI_SW};')
    e(f'mov.b32 hno, {HI_NONE};')
    e(f'mov.b32 hti, {HI_TIMG};')
    e(f'mov.b32 hz, {HI_TIMG if znone else HI_SW};')
    e(f'mov.b32 iN16, {idesc(128, 16, 0, 0)};')
    e(f'mov.b32 iEG, {idesc(128, EGN, 0, 0)};')
    e(f'mov.b32 iDX, {idesc(128, DEP, 0, 0)};')
    e(f'mov.b32 iDQ, {idesc(128, D, 0, 1)};')
    e(f'mov.b32 iT, {idesc(128, 16, 1, 1)};')
    e(f'mov.b32 iW, {idesc(128, DEP, 1, 1)};')
    # ---- products of key j ----
    # dQ += dS Kexp
    e(f'add.u32 ra, {tg}, {G_DA};')
    p.mma_ts('pe', tm_dq, 'ra', loKc_mn, 'hsw', 'iDQ', 'pacc')
    # T = de' I + (r dZ) W'^T
    for s in range(DEP // 16):
        e(f'add.u32 rd, {tg}, {G_DX + 16 * s};')
        e(f'add.u32 ra, {ldc}, {2 * s if DE >= 16 else 0};')
        p.mma_ss('pe', 'rd', 'ra', 'hsw', loI, 'hno', 'iN16', 'pz')
    e(f'add.u32 rd, {tg}, {G_DX};')
    for s in range(EGN // 16):
        e(f'add.u32 ra, {tg}, {G_S + 8 * s};')
        e(f'add.u32 rb, {loWdx}, {2 * s * DEP};')
        p.mma_ts('pe', 'rd', 'ra', 'rb', 'hno', 'iDX', 'pt')
    # dK^T, dV^T: contraction over the query rows (k-steps >= KR skipped)
    for which, (loX, img, col) in enumerate(((loQm, loS, G_T), (loDOm, loA, G_T + 16))):
        e(f'add.u32 rd, {tg}, {col};')
        for s in range(8):
            e(f'add.u32 ra, {loX}, {128 * s};')
            e(f'add.u32 rb, {img}, {16 * s};')
            if s == 0:
                p.mma_ss('pe', 'rd', 'ra', 'hsw', 'rb', 'hti', 'iT', 'pz')
            else:
                e(f'setp.le.u32 pk, {KR}, {s};')
                e(f'@pk bra LT{which};')
                p.mma_ss('pk', 'rd', 'ra', 'hsw', 'rb', 'hti', 'iT', 'pt')
        e(f'LT{which}:')
    # weight-gradient accumulators
    zstep = 16 if znone else 128
    for wi, (dcol, win) in enumerate(((tm_w1, we), (tm_w2, wd))):
        for s in range(8):
            e(f'add.u32 ra, {loZ}, {zstep * s};')
            e(f'add.u32 rb, {win}, {128 * s};')
            if s == 0:
                p.mma_ss('pe', dcol, 'ra', 'hz', 'rb', 'hsw', 'iW', 'paccw')
            else:
                e(f'setp.le.u32 pk, {KR}, {s};')
                e(f'@pk bra LW{wi};')
                p.mma_ss('pk', dcol, 'ra', 'hz', 'rb', 'hsw', 'iW', 'pt')
        e(f'LW{wi}:')
    # ---- inputs of the group's next key ----
    e('@!pn bra LCOMMIT;')
    for col, loX, loB in ((G_S, loQ, loKn), (G_DA, loDO, loVn)):
        e(f'add.u32 rd, {tg}, {col};')
        for s in range(DKS):
            e(f'add.u32 ra, {loX}, {(s >> 2) * 1024 + (s & 3) * 2};')
            e(f'add.u32 rb, {loB}, {(s >> 2) * 128 + (s & 3) * 2};')
            p.mma_ss('pn', 'rd', 'ra', 'hsw', 'rb', 'hsw', 'iN16', 'pz' if s == 0 else 'pt')
    e(f'add.u32 rd, {tg}, {G_EG};')
    for s in range(2 * EK):                               # W' = hi + lo
        e(f'add.u32 ra, {len_}, {2 * (s % EK)};')
        e(f'add.u32 rb, {loWeg}, {2 * s * EGN};')
        p.mma_ss('pn', 'rd', 'ra', 'hsw', 'rb', 'hno', 'iEG', 'pz' if s == 0 else 'pt')
    e(f'add.u32 rd, {tg}, {G_HX};')
    for s in range(EK):
        e(f'add.u32 ra, {ldn}, {2 * s};')
        e(f'add.u32 rb, {loWhx}, {32 * s};')
        p.mma_ss('pn', 'rd', 'ra', 'hsw', 'rb', 'hno', 'iN16', 'pz' if s == 0 else 'pt')
    e('LCOMMIT:')
    e(f'tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [{bar}];')
    e('LDONE:')
    n_mma = sum(1 for l in p.lines if 'tcgen05.mma' in l)
    args = ', '.join(f'uint32_t {n}' for n, _ in p.ops)
    body = '\n'.join(f'      "{l}\\n\\t"' for l in p.lines)
    cons = ', '.join(f'"r"({n})' for n, _ in p.ops)
    return f'''// {name}: h = {H}, dk = {DK}, d_e = {DE}: {n_mma} tcgen05.mma + 1 commit
__device__ __forceinline__ void wide_bwd_program_{name}({args}) {{
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
print('// The tcgen05.mma program of one backward handshake of wide_bwd.cu as ONE inline-asm statement per instantiation')
print('// (operands: warp-uniform 32-bit values; every lane of the converged issuer warp executes it, one elected lane issues).')
print('#pragma once')
print('#include <stdint.h>')
print()
print('namespace egt {')
print()
print(gen('c5', 16, 8, 32, False))
print(gen('c3', 8, 12, 8, False))
print(gen('c1', 8, 8, 64, True))
print('}  // namespace egt')
