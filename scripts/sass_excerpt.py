"""profiles/r2_sass_*.txt: per hot kernel of libctcb.so its resource usage, opcode histogram and a SASS excerpt around the
instruction class that characterises it (cuobjdump, no GPU needed).  usage: python scripts/sass_excerpt.py"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gluon_e2e_asr_b200", "libctcb.so")
HOT = [  # (file tag, mangled-name regex, opcode to centre the excerpt on)
    ("k_walk_P2_NW2_hist_fused", r"k_walkILi2ELi2ELb1ELb1E", "DFMA"),
    ("k_grad_VEC2_CH4_XQ1", r"k_gradILi2ELi4ELi1ELi0E", "REDUX"),
    ("k_grad2_VEC2_CH4_occ", r"k_grad2ILi2ELi4ELi1E", "DMUL"),
    ("k_emit_staged", r"k_emitILi4ELin1E", "UBLKCP"),
    ("k_grad_staged_CH8", r"k_gradILi4ELi8ELin1ELi0E", "UBLKCP"),
    ("k_walk_P2_NW3_unfused", r"k_walkILi2ELi3ELb1ELb0E", "UBLKCP"),
    ("k_meet_P4", r"k_meetILi4E", "UBLKCP"),
    ("k_proj_emit_tcgen05", r"k_proj_emit", "UTCHMMA"),
]
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for tag, rx, centre in HOT:
    names = [n for n in funcs if re.search(rx, n)]
    if not names:
        print("not found:", rx); continue
    name = names[0]; ins = funcs[name]
    usage = ""
    lines = res.split("\n")
    for i, l in enumerate(lines):
        if name in l and i + 1 < len(lines):
            usage = lines[i + 1].strip()
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    idx = [i for i, (_, t) in enumerate(ins) if centre in t]
    mid = idx[len(idx) // 2] if idx else 0
    lo, hi = max(0, mid - 30), min(len(ins), mid + 30)
    with open(os.path.join(ROOT, "profiles", "r2_sass_%s.txt" % tag), "w") as f:
        f.write("%s\n%s\n%d SASS instructions (%d bytes)\n\nopcode histogram (static):\n" % (name, usage, len(ins), 16 * len(ins)))
        f.write("  " + "  ".join("%s %d" % kv for kv in ops.most_common(28)) + "\n")
        f.write("\nexcerpt around %s (%d of them):\n" % (centre, len(idx)))
        for a, t in ins[lo:hi]:
            f.write("  /*%05x*/  %s\n" % (a, t))
    print(tag, usage.split(" STACK")[0], len(ins), "instr;", centre, len(idx))
