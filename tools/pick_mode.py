"""Reads gpurun_out/probe_conv.json and prints shell exports selecting the kernel variant for the
rest of a GPU session: the first tcgen05 descriptor mode that passes every case, else the SIMT path."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
s = json.load(open(os.path.join(REPO, 'gpurun_out', 'probe_conv.json')))


def all_ok(v):
    c = s.get(v, {}).get('cases')
    return bool(c) and all(x.get('ok') for x in c.values()) and s[v].get('device_flag', 1) == 0


for v in ['v2', 'v1']:
    if all_ok(v):
        print('export TTSB_CONV_IMPL=tc TTSB_DESC_MODE=0 TTSB_TC_VERSION=%s' % v[1])
        sys.exit(0)
print('export TTSB_CONV_IMPL=simt')
