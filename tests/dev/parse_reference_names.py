"""Extract the expression names the reference's planners pass as `name=` arguments and how each is added
(add_dynamics / add_expression_to_horizon / add_expression / add_constraint / add_cost, and minimize vs subject_to where
the call says so):
  /root/reference/src/hippopt/turnkey_planners/humanoid_kinodynamic/planner.py -> tests/golden/reference_expression_names.json
  /root/reference/src/hippopt/turnkey_planners/humanoid_pose_finder/planner.py -> tests/golden/reference_pose_expression_names.json
Run in the build container (the reference is not on the GPU box).
"""
import ast
import json
import os
import sys

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")
JOBS = [("/root/reference/src/hippopt/turnkey_planners/humanoid_kinodynamic/planner.py",
         os.path.join(GOLD, "reference_expression_names.json")),
        ("/root/reference/src/hippopt/turnkey_planners/humanoid_pose_finder/planner.py",
         os.path.join(GOLD, "reference_pose_expression_names.json"))]
CALLS = ("add_dynamics", "add_expression_to_horizon", "add_expression", "add_constraint", "add_cost")


def render(node) -> str:
    """name= expression -> pattern: string literals verbatim, `<x>.name()` -> "<point>.<x's last attribute>",
    a bare identifier (foot_name) -> "<identifier>"."""
    if isinstance(node, ast.Constant) and isinstance(node.value, str):
        return node.value
    if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
        return render(node.left) + render(node.right)
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "name":
        inner = node.func.value
        return "<point>." + (inner.attr if isinstance(inner, ast.Attribute) else ast.unparse(inner))
    if isinstance(node, ast.Name):
        return f"<{node.id}>"
    return ast.unparse(node)


def parse(REF, OUT):
    tree = ast.parse(open(REF).read())
    out = []
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)):
            continue
        if node.func.attr not in CALLS:
            continue
        kw = {k.arg: k.value for k in node.keywords}
        if "name" not in kw:
            continue
        mode = ast.unparse(kw["mode"]) if "mode" in kw else "default(subject_to)"
        out.append({"call": node.func.attr, "name": render(kw["name"]), "mode": mode, "line": node.lineno,
                    "apply_to_first_elements": ast.unparse(kw["apply_to_first_elements"])
                    if "apply_to_first_elements" in kw else None})
    out.sort(key=lambda d: d["line"])
    json.dump({"source": REF, "expressions": out}, open(OUT, "w"), indent=1)
    print(f"{len(out)} named expressions -> {OUT}")


def main():
    for ref, out in JOBS:
        parse(ref, out)


if __name__ == "__main__":
    sys.exit(main())
