# pass 23 (1 GPU): the whole GPU suite with FastVGICP in; smoke
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
