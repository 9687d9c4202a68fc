"""Instruction / stall-sample share per CUDA source line of an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
cur = None; hdr = None; out = []
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and r[2] == '-':
        try: out.append((cur, int(r[0]), r[1].strip()[:100], int(r[6] or 0), int(r[7] or 0)))
        except ValueError: pass
ti = sum(o[4] for o in out); ts = sum(o[3] for o in out)
print('total warp-inst', ti, 'samples', ts)
for f, l, src, s, i in sorted(out, key=lambda o: -(o[4] / ti + o[3] / ts))[:top]:
    print(f'{f[:16]:16s} {l:4d} inst {i / ti * 100:5.2f}% samp {s / ts * 100:5.2f}%  {src}')
