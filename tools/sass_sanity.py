"""Static sanity scan of the fused kernels' SASS (ptxas 12.9 was seen to miscompile one instantiation of k_resid7: a
shared-memory address built on the stack pointer R1 / on an unrelated 4-byte-scaled register).  Flags, per kernel:
  * any use of R1 in a kernel without a stack frame;
  * an 8-byte-scaled address (IMAD Rd, Rx, 0x8, Rb) whose base Rb was last defined as a 4-byte-scaled thread offset.
usage: python tools/sass_sanity.py hifiles-solver_b200/build/hf_fused.cu.o"""
import re, subprocess, sys, tempfile, os

def main(obj):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    stack = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:\d+ STACK:(\d+)", res):
        stack[m.group(1)] = int(m.group(2))
    fn, defs, bad = None, {}, 0
    for l in sass:
        m = re.match(r"\.text\.(\S+):", l)
        if m:
            fn, defs = m.group(1), {}
            continue
        m = re.search(r"/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?(\S+)\s+(.*);", l)
        if not m or fn is None:
            continue
        pc, op, args = m.group(1), m.group(3), m.group(4)
        a = [x.strip() for x in args.split(",")]
        if stack.get(fn, 0) == 0 and re.search(r"(?<![A-Za-z0-9])R1(?![0-9])", args) and "c[0x0][0x37c]" not in args:
            print("R1 used without a stack frame:", fn, pc, l.strip()[:100]); bad += 1
        if op.startswith("IMAD") and len(a) == 4 and a[2] == "0x4" and re.match(r"R\d+$", a[0]):
            defs[a[0]] = pc
            continue
        if op.startswith("IMAD") and len(a) == 4 and a[2] == "0x8" and a[3] in defs:
            print("8-byte address on a 4-byte-scaled base:", fn, pc, l.strip()[:100]); bad += 1
        d0 = re.match(r"(R\d+)", a[0]) if a else None
        if d0 and d0.group(1) in defs and not op.startswith(("ST", "LDGSTS", "BRA", "ISETP", "RED", "ATOM")):
            del defs[d0.group(1)]
    print("kernels scanned: %d, findings: %d" % (len(stack), bad))
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
