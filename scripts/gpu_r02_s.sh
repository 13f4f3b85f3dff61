mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Netlist | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Netlist | tail -2
