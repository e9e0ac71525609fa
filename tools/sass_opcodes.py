#!/usr/bin/env python
"""SASS opcode mix of the hot kernels, static (cuobjdump of the built library) and dynamic (instructions executed, from
the source page of an `ncu --set full --import-source on` report):

    python tools/sass_opcodes.py gpurun_out/prof_r02a.ncu-rep > profiles/r02_sass_opcodes.md

The dynamic table is what the issue slots were spent on; the static one is the proof of which instructions the
compiler emitted for sm_100a (LDGSTS = cp.async, MUFU.RCP64H = the reciprocal seed, no HMMA/UTCMMA: there is no dense
contraction on this path)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "montecarlo_simulation_b200", "libmc2d.so")
KERNELS = [("k_sweep_p<LJ, G=4>", "_Z9k_sweep_pILi0ELi4ELb0EEv9SweepArgsi", r"k_sweep_p<\(int\)0, \(int\)4, \(bool\)0>"),
           ("k_energy_p<LJ, G=4>", None, r"k_energy_p<\(int\)0, \(int\)4>"),
           ("k_rebin4<4>", None, r"k_rebin4<\(int\)4")]
GROUPS = [("fp64 (DFMA DADD DMUL DSETP)", ("DFMA", "DADD", "DMUL", "DSETP")), ("MUFU", ("MUFU",)),
          ("integer / logic (IMAD IADD3 LOP3 SHF LEA ISETP SEL VIADD ...)",
           ("IMAD", "IADD3", "IADD", "LOP3", "SHF", "LEA", "ISETP", "SEL", "VIADD", "VIMNMX", "IDP", "PRMT", "I2F", "F2I",
            "FSEL", "PLOP3", "MOV", "CS2R", "S2R", "POPC", "FLO", "VIADDMNMX", "IABS", "HFMA2", "VIMNMX3", "I2FP", "F2F")),
          ("shared memory (LDS STS)", ("LDS", "STS")), ("cp.async (LDGSTS LDGDEPBAR DEPBAR)", ("LDGSTS", "LDGDEPBAR", "DEPBAR")),
          ("global / local memory (LDG STG LDL STL ATOM RED CCTL MEMBAR)", ("LDG", "STG", "LDL", "STL", "ATOM", "ATOMG", "RED", "CCTL", "MEMBAR", "ERRBAR", "CGAERRBAR")),
          ("shuffle / vote (SHFL VOTE MATCH)", ("SHFL", "VOTE", "MATCH")),
          ("control (BRA BSSY BSYNC WARPSYNC BAR EXIT ...)", ("BRA", "BSSY", "BSYNC", "WARPSYNC", "BAR", "EXIT", "NOP", "BREAK", "CALL", "RET", "ENDCOLLECTIVE", "NANOSLEEP", "YIELD", "BMOV")),
          ("constant / uniform (LDC LDCU UMOV R2UR U*)", ("LDC", "LDCU", "UMOV", "R2UR", "UIADD3", "UIMAD", "UISETP", "USEL", "USHF", "ULOP3", "ULEA", "UPLOP3"))]


def opcode(text):
    t = text.strip()
    parts = t.split()
    if parts and parts[0].startswith("@"):
        parts = parts[1:]
    return parts[0].split(".")[0] if parts else ""


def table(counter, title, out):
    tot = sum(counter.values())
    if tot == 0:
        out.append("\n%s: not in this report" % title)
        return
    out.append("\n%s (%d)\n\n| group | share | top opcodes |\n|---|---:|---|" % (title, tot))
    seen = set()
    for name, ops in GROUPS:
        sub = {o: counter[o] for o in ops if counter.get(o)}
        seen.update(sub)
        n = sum(sub.values())
        top = ", ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in sorted(sub.items(), key=lambda kv: -kv[1])[:5])
        out.append("| %s | %.1f%% | %s |" % (name, 100.0 * n / tot, top))
    rest = {o: c for o, c in counter.items() if o not in seen}
    if rest:
        out.append("| other | %.1f%% | %s |" % (100.0 * sum(rest.values()) / tot,
                                               ", ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in sorted(rest.items(), key=lambda kv: -kv[1])[:6])))


def static_mix(symbol_regex):
    names = subprocess.check_output(["cuobjdump", "-sass", LIB], stderr=subprocess.DEVNULL).decode()
    cur, counts = None, collections.Counter()
    for line in names.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.search(symbol_regex, cur):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
            if m:
                counts[opcode(m.group(1))] += 1
    return counts


def dynamic_mix(rep, kernel_regex):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                                   "regex:" + kernel_regex.split("<")[0]], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(raw.splitlines()))
    counts, hdr, take, done = collections.Counter(), None, False, False
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            if done:
                break
            take = re.search(kernel_regex, r[1]) is not None
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if take and hdr:
            counts[opcode(r[hdr.index("Source")])] += int(r[hdr.index("Instructions Executed")])
            done = True
    return counts


def main():
    rep = sys.argv[1] if len(sys.argv) > 1 else None
    reps = sys.argv[1:]  # several reports: the first one that holds the kernel is used
    out = ["# SASS opcode mix of the hot kernels (sm_100a build of libmc2d.so)\n",
           "Static = instructions in the binary (`cuobjdump -sass`); dynamic = warp instructions executed in one launch "
           "(config #4, one colour pass / one reduction / one re-bin), from the source page of `%s`." % (os.path.basename(rep) if rep else "-")]
    static_syms = {"k_sweep_p<LJ, G=4>": r"k_sweep_pILi0ELi4ELb0E", "k_energy_p<LJ, G=4>": r"k_energy_pILi0ELi4E", "k_rebin4<4>": r"k_rebin4ILi4ELb0E"}
    for title, _, kre in KERNELS:
        out.append("\n## `%s`" % title)
        table(static_mix(static_syms[title]), "static", out)
        for rp in reps:
            dyn = dynamic_mix(rp, kre)
            if sum(dyn.values()):
                table(dyn, "dynamic (instructions executed, %s)" % os.path.basename(rp), out)
                break
    print("\n".join(out))


if __name__ == "__main__":
    main()
