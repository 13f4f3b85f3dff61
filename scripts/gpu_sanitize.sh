timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_asm_tran.py 2>&1 | grep -v Netlist | tail -4
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_asm_tran.py 2>&1 | grep -v Netlist | tail -6
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python scripts/sanitize_asm_tran.py 2>&1 | grep -v Netlist | tail -4
