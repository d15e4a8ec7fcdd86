""" SASS-level view of one ncu report: stall reasons over all warp samples, the opcode mix by executed instructions and by
samples, and which opcodes sit behind the long-scoreboard waits.

    python tools/ncu_sass_mix.py gpurun_out/prof_cfg5_r2d.ncu-rep > profiles/r2d_cfg5_wide_kernel_sass_mix.txt
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    print(rows[0][1] if len(rows[0]) > 1 else '')
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    S, I = ix['# Samples'], ix['Instructions Executed']
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[S]) for r in data)
    toti = sum(int(r[I]) for r in data)
    print('static instructions %d (executed at least once: %d), warp-level instructions executed %d, samples %d'
          % (len(data), sum(1 for r in data if int(r[I]) > 0), toti, tot))
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print('stall reasons (share of all samples):')
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
        print('  %-24s %5.1f%%' % (h[6:], 100 * v / tot))

    def op(r):
        m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[1])
        return m.group(2) if m else '?'
    opi, ops, lsb = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        o = op(r)
        opi[o] += int(r[I]); ops[o] += int(r[S]); lsb[o] += int(r[ix['stall_long_sb']])
    print('opcode mix (share of executed instructions / share of samples):')
    for o, v in opi.most_common(30):
        print('  %-10s %5.1f%%  %5.1f%%' % (o, 100 * v / toti, 100 * ops[o] / tot))
    tl = max(1, sum(lsb.values()))
    print('long-scoreboard waits by waiting opcode: ' + ', '.join('%s %.1f%%' % (k, 100 * v / tl) for k, v in lsb.most_common(10)))


if __name__ == '__main__':
    main()
