"""Stall-reason totals, instruction mix and the most-stalled SASS lines of one kernel from an .ncu-rep that was captured with
`--set full --import-source on` (ncu -i ... --page source --csv).

    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep [units_per_launch] [launch_index]   # e.g. 5000 tiles -> per-tile counts
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    sel = ['--launch-skip', sys.argv[3], '--launch-count', '1'] if len(sys.argv) > 3 else []
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print('kernel:', rows[0][1])
    hdr = rows[1]
    seen, data = set(), []
    for r in rows[2:]:            # (a report with both SASS and source views lists every address twice)
        if len(r) >= len(hdr) - 1 and r[0].startswith('0x') and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    samples = sum(int(r[ix['# Samples']]) for r in data)
    instr = sum(int(r[ix['Instructions Executed']]) for r in data)
    print('SASS instructions: %d   warp-level instructions executed: %.4e%s   stall samples: %d' % (
        len(data), instr, ('  (%.0f per unit)' % (instr / units)) if units else '', samples))
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    print('\nwarp stall samples by reason:')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print('  %-26s %6.1f %%' % (k, 100.0 * v / samples))
    ops = collections.Counter()
    for r in data:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
        ops[m.group(2).split('.')[0] if m else '?'] += int(r[ix['Instructions Executed']])
    print('\nexecuted warp-level instructions by opcode:')
    for k, v in ops.most_common(24):
        print('  %-12s %5.1f %%%s' % (k, 100.0 * v / instr, ('   %8.0f per unit' % (v / units)) if units else ''))
    print('\nmost-stalled instructions (samples, share, executed, SASS, top two reasons):')
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:25]:
        top = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
        print('  %6d %5.2f %%  ex=%9s  %-56s %s' % (int(r[ix['# Samples']]), 100.0 * int(r[ix['# Samples']]) / samples,
                                                  r[ix['Instructions Executed']], r[ix['Source']].strip()[:56],
                                                  ', '.join('%s %d' % (s, c) for c, s in top)))


if __name__ == '__main__':
    main()
