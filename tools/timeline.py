"""Reads the text file written by a trainer run under NNCF_TIMELINE=<file> (global-timer stamps taken inside the
gather and score kernels) and prints the in-situ step timeline: how long each kernel really runs inside the
programmatic-dependent-launch chain and where the gaps are.  usage: python tools/timeline.py file [skip]"""
import sys
import numpy as np

rows = [list(map(int, l.split())) for l in open(sys.argv[1]) if l.strip() and not l.startswith('#')]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 20
a = np.array(rows[skip:], dtype=np.float64)
a = a[(a[:, 0] < 2 ** 63) & (a[:, 5] > 0)]
g0, gw, ge, s0, sw, se, sl = [a[:, i] for i in range(7)]
def med(x): return float(np.median(x)) / 1e3
print("steps analysed: %d" % len(a))
print("gather: first CTA resident -> past griddepcontrol.wait  %.2f us" % med(gw - g0))
print("gather: past wait -> last CTA done                      %.2f us" % med(ge - gw))
print("score : first CTA resident -> past wait                 %.2f us (prologue overlapped with the gather)" % med(sw - s0))
print("score : gather end -> first CTA past wait               %.2f us" % med(sw - ge))
print("score : past wait -> last CTA leaves tile loop          %.2f us" % med(sl - sw))
print("score : tile loop end -> last CTA done (drain + update) %.2f us" % med(se - sl))
print("score : first -> LAST CTA past wait (residency stagger)   %.2f us" % med(a[:, 8] - sw))
print("score : CTA 0 past wait -> end: %.0f cycles in %.2f us = %.0f MHz effective SM clock" % (np.median(a[:, 9]), med(a[:, 10]), np.median(a[:, 9]) / max(np.median(a[:, 10]), 1) * 1e3))
print("next gather past wait - score end                       %.2f us" % med(gw[1:] - se[:-1]))
print("step period (gather wait -> next gather wait)           %.2f us" % med(gw[1:] - gw[:-1]))
