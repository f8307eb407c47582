"""Static evidence from the built library (runs without a GPU): per-kernel SASS mnemonic counts that show which
kernels use the Blackwell paths (tcgen05 MMAs, TMEM loads, TMA), and the ptxas resource table (registers, spills,
shared memory) of a verbose rebuild.

    python tools/sass_evidence.py [--ptxas] > profiles/r2_sass_mnemonics.txt

Mnemonics (see /opt/skills/guides/B200_PROFILING.md): UTCHMMA = tcgen05.mma (bf16), LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc / dealloc, UTMALDG = TMA tensor load (cp.async.bulk.tensor),
UBLKCP = cp.async.bulk (non-tensor bulk copy), LDGSTS = cp.async, SYNCS = mbarrier ops, MUFU.EX2 = exp2,
RED / REDG / ATOMG = global reductions.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'tartangan_b200', 'lib', 'libttg_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'MUFU.EX2', 'REDG', 'ATOMG',
        'HMMA', 'STL', 'LDL']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return [re.sub(r'\(.*$', '', o).replace('void ', '') for o in out]


def sass_table():
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        cur['_n'] += 1
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k == 'HMMA' and op.startswith('HMMA')):
                cur[k] += 1
    names = list(per)
    pretty = demangle(names)
    total = collections.Counter()
    print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(names)} kernels')
    print('# kernels that use tcgen05 / TMEM / TMA / bulk copies (instantiations of one template merged: count x name)')
    merged = collections.OrderedDict()
    for n, p in zip(names, pretty):
        c = per[n]
        total.update(c)
        base = re.sub(r'<.*$', '', p)
        m = merged.setdefault(base, [0, collections.Counter()])
        m[0] += 1
        m[1].update(c)
    cols = [k for k in KEYS if k not in ('STL', 'LDL', 'HMMA')]
    print(f'{"kernel (template)":58s} {"inst":>4s} {"SASS":>8s} ' + ' '.join(f'{k:>10s}' for k in cols))
    for base, (cnt, c) in merged.items():
        if not any(c[k] for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP')):
            continue
        print(f'{base[:58]:58s} {cnt:4d} {c["_n"]:8d} ' + ' '.join(f'{c[k]:10d}' for k in cols))
    print()
    print('# whole library: ' + ', '.join(f'{k} {total[k]}' for k in KEYS) + f', instructions {total["_n"]}')
    print(f'# mma.sync-style HMMA instructions (the pre-Blackwell tensor-core path): {total["HMMA"]}')
    print(f'# kernels without any tcgen05 / TMA instruction (streaming, reduction, optimiser, loss kernels): '
          f'{sum(1 for n in names if not any(per[n][k] for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP")))}')


def ptxas_table():
    """Verbose rebuild into a scratch variant (the shipped library is left alone)."""
    env = dict(os.environ, TTG_BUILD_VARIANT='ptxasv')
    r = subprocess.run([sys.executable, '-m', 'tartangan_b200.build', '--force', '-v'], cwd=ROOT, env=env,
                       capture_output=True, text=True)
    log = r.stdout + r.stderr
    rows, name = [], None
    for line in log.split('\n'):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            name = m.group(1)
            spill = None
            continue
        m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
        if m and name:
            spill = tuple(int(x) for x in m.groups())
            continue
        m = re.search(r'Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?', line)
        if m and name:
            rows.append((name, int(m.group(1)), spill or (0, 0, 0), int(m.group(3) or 0)))
            name = None
    import shutil
    shutil.rmtree(os.path.join(ROOT, 'tartangan_b200', 'build_ptxasv'), ignore_errors=True)
    try:
        os.remove(LIB[:-3] + '_ptxasv.so')
    except OSError:
        pass
    pretty = demangle([r[0] for r in rows])
    print()
    print(f'# nvcc -Xptxas=-v (sm_100a): {len(rows)} entry functions; registers / static smem bytes / spills')
    hist = collections.Counter()
    for (n, regs, sp, smem), p in zip(rows, pretty):
        hist[min(255, (regs + 31) // 32 * 32)] += 1
    print('# register histogram (<= bucket: kernels): ' + ', '.join(f'<={k}: {v}' for k, v in sorted(hist.items())))
    spilled = [(p, regs, sp) for (n, regs, sp, smem), p in zip(rows, pretty) if sp[1] or sp[2]]
    print(f'# kernels with register spills: {len(spilled)} of {len(rows)}')
    for p, regs, sp in spilled:
        print(f'  {p[:110]:110s} regs {regs:3d} stack {sp[0]:4d} B, spill stores {sp[1]:4d} B, loads {sp[2]:4d} B')
    print('# tcgen05 / TMA kernels (largest register counts first)')
    hot = [(regs, p, sp) for (n, regs, sp, smem), p in zip(rows, pretty)
           if re.search(r'conv_tc|wgrad_tc|attn_tc|attention_tc|conv_wgrad', p)]
    for regs, p, sp in sorted(hot, reverse=True)[:40]:
        print(f'  {p[:120]:120s} regs {regs:3d} spill loads {sp[2]} B')


if __name__ == '__main__':
    sass_table()
    if '--ptxas' in sys.argv:
        ptxas_table()
