"""Instruction mix of the kernels in a `cuobjdump -sass` dump whose (mangled) name matches a regex:
    cuobjdump -sass simplediffeq.jl_b200/build/sde_sys_lorenz.cu.o > /tmp/l.sass
    python tools/sass_mix.py /tmp/l.sass 'adaptive_kernel.*Vern9MethodIS1_dEELi0ELb1ELb1'
Development aid (static counts over the whole kernel, cold paths included)."""
import collections
import re
import sys


def main():
    text = open(sys.argv[1]).read()
    pat = re.compile(sys.argv[2])
    for blk in text.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not pat.search(name):
            continue
        ops = collections.Counter()
        n = 0
        for ln in blk.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m:
                ops[m.group(1)] += 1
                n += 1
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
        print("%s\n  %d instructions, %d FP64-pipe (DFMA/DMUL/DADD/DSETP/MUFU); top: %s" % (
            name, n, fp64, " ".join("%s:%d" % kv for kv in ops.most_common(14))))


if __name__ == "__main__":
    main()
