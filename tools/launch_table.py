"""Join an ncu launch list (gpu__time_duration) of tools/net_bench.py with the plan text: python tools/launch_table.py CSV MODEL"""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rs_face_detection_tflite_b200 as fdl
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
net = fdl.Net('models/%s.tflite' % sys.argv[2], -1)
steps = [l for l in net.describe().splitlines() if l.startswith('#')]
n = len(steps)
last = rows[-n:]
tot = sum(float(r['Metric Value']) for r in last)
print('steps', n, 'total %.1f us' % (tot / 1e3))
for s, r in zip(steps, last):
    t = float(r['Metric Value']) / 1e3
    k = 'TC ' if 'blaze_block_tc' in r['Kernel Name'] else ('elt' if 'elementwise' in r['Kernel Name'] else 'v1 ')
    print('%8.1f us %5.1f%% %s grid=%-14s %s' % (t, 100 * t * 1e3 / tot, k, r['Grid Size'], s[:110]))
