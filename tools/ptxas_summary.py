""" Per-kernel resource table from the `-Xptxas -v` log the build writes (build/ptxas.log):
    python tools/ptxas_summary.py > profiles/r2_ptxas_summary.txt
Static evidence (no GPU needed): registers, spills, stack, static shared memory of every sm_100a kernel in
libpinn_b200.so — the thing to check before spending GPU time on a change. """
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    try:
        out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True, check=True).stdout
        return out.splitlines()
    except (OSError, subprocess.CalledProcessError):
        return names


def main():
    log = open(os.path.join(ROOT, 'build', 'ptxas.log')).read()
    rows = []
    pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                     r"(\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?"
                     r"(?:, (\d+) bytes smem)?", re.S)
    for m in pat.finditer(log):
        rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(8) or 0)))
    names = demangle([r[0] for r in rows])
    short = []
    for n in names:
        n = re.sub(r'\(pinn::DevPlan.*', '', n).replace('void pinn::', '').replace('(int)', '').replace('(bool)', '')
        short.append(n)
    print('%d kernels; columns: registers / stack bytes / spill-store bytes / spill-load bytes / static smem bytes' % len(rows))
    print('template arguments: step_kernel<NF, NS, GMEM, MAXT, JF, GEN>, multi_step_kernel<NF, NS, MAXT, JF>, '
          'wide_step_kernel<NF, NS, THREADS>, small_step_kernel<NF, NS>')
    for n, r in sorted(zip(short, rows)):
        print('%-58s regs %3d  stack %5d  spill st/ld %5d/%5d  smem %5d' % (n, r[1], r[2], r[3], r[4], r[5]))
    spilled = [n for n, r in zip(short, rows) if r[3] or r[4]]
    print('kernels with register spills: %d%s' % (len(spilled), (' (' + ', '.join(sorted(spilled)) + ')') if spilled else ''))


if __name__ == '__main__':
    sys.exit(main())
