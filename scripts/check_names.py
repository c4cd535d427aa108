#!/usr/bin/env python
"""Poor man's pyflakes (no linter in the image): report names a function reads as globals that the module never defines."""
import ast, builtins, symtable, sys

def check(path):
    src = open(path).read()
    tree = ast.parse(src)
    top = symtable.symtable(src, path, "exec")
    module_names = set(top.get_identifiers()) | set(dir(builtins))
    bad = []
    def walk(t):
        for c in t.get_children():
            for sym in c.get_symbols():
                if sym.is_global() and sym.is_referenced() and sym.get_name() not in module_names:
                    bad.append((c.get_name(), sym.get_name()))
                if sym.is_free():
                    pass
            walk(c)
    walk(top)
    # local names read but never bound anywhere in the enclosing function chain are reported by symtable as globals,
    # which the loop above already covers
    return bad

rc = 0
for p in sys.argv[1:]:
    for fn, name in check(p):
        print(f"{p}: function {fn}: undefined name {name}")
        rc = 1
sys.exit(rc)
