#!/usr/bin/env python3
"""Records the reference's StepDemonstrator (utils/demonstrator.py) on a scripted call sequence ->
tests/golden/step_demonstrator.json.  Runs only in the build container (imports /root/reference)."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_demonstrator", "/root/reference/pybullet_multigoal_gym/utils/demonstrator.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

demos = [list(range(i + 1)) for i in range(4)]
script = [("next",)] * 3 + [("manual_reset", 2)] + [("next",)] * 6 + [("reset_last", 3)] + [("next",)] * 7 + [("manual_reset", None)] + [("next",)] * 3
out = []
for stick in (True, False):
    d = mod.StepDemonstrator(demos, stick_with_final_goal=stick)
    trace = []
    for op in script:
        if op[0] == "next":
            trace.append(["next", d.get_next_goal(), bool(d.final), d.current_goal, d.demon_ind])
        elif op[0] == "manual_reset":
            d.manual_reset(op[1])
            trace.append(["manual_reset", op[1], bool(d.final), d.current_goal, d.demon_ind, d.current_final_goal])
        else:
            d.reset_with_the_last_sub_goal_index(op[1])
            trace.append(["reset_last", op[1], bool(d.final), d.current_goal, d.demon_ind, d.current_final_goal])
    out.append({"stick_with_final_goal": stick, "trace": trace})
json.dump({"demonstrations": demos, "runs": out}, open(os.path.join(ROOT, "tests", "golden", "step_demonstrator.json"), "w"), indent=1)
print("wrote tests/golden/step_demonstrator.json")
