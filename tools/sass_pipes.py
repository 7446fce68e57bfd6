"""Static instruction-mix of a kernel's SASS: how many instructions go to the ALU pipe, the FMA pipe, LSU, branch.
Usage: python tools/sass_pipes.py <lib.so> <kernel-name-substring> [--dump]"""
import re
import subprocess
import sys

ALU = ("LOP3", "SEL", "ISETP", "FSETP", "IADD3", "SHF", "VIMNMX", "FMNMX", "PRMT", "LEA", "P2R", "R2P", "FSEL", "PLOP3", "IABS", "FLO", "POPC", "BREV", "VIADD", "MOV", "CS2R", "I2FP", "F2FP", "FCHK", "VIADDMNMX")
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2")
LSU = ("LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "LDCU", "ATOM", "RED")
BR = ("BRA", "BSSY", "BSYNC", "BREAK", "EXIT", "RET", "CALL", "WARPSYNC", "NOP", "BAR")


def main():
    lib, name = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, keep, lines = None, False, []
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            keep = name in m.group(1)
            if keep:
                lines.append("## " + m.group(1))
            continue
        if keep:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
            if m:
                lines.append(m.group(1) + " " + m.group(2).strip())
    if "--dump" in sys.argv:
        print("\n".join(lines))
        return
    cnt = {"alu": 0, "fma": 0, "lsu": 0, "br": 0, "other": 0}
    for l in lines:
        if l.startswith("##"):
            if sum(cnt.values()):
                print(cnt)
            print(l)
            cnt = {k: 0 for k in cnt}
            continue
        op = l.split()[1] if l.split()[1][0] != "@" else l.split()[2]
        op = op.split(".")[0]
        if op in ALU: cnt["alu"] += 1
        elif op in FMA: cnt["fma"] += 1
        elif op in LSU: cnt["lsu"] += 1
        elif op in BR: cnt["br"] += 1
        else: cnt["other"] += 1; print("  ?", op)
    print(cnt)


if __name__ == "__main__":
    main()
