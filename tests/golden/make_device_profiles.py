#!/usr/bin/env python
"""Convert the reference's device noise profiles to a JSON fixture.

Input : /root/reference/resources/simulator-devices.edn (data file read by
        adapter/backend/hardware_simulator.clj:40-45)
Output: tests/golden/device_profiles.json  — for every device: id, name, num-qubits,
        native gates and the :noise-model map ({:gate-noise {gate {...}} :readout-error {...}})
        exactly as in the EDN (keywords spelled ":kw").  Coupling maps are dropped (not used on the
        simulation path).  Runs only in the build container; the JSON is committed.
"""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import edn
from make_tutorial_golden import jsonable

SRC = "/root/reference/resources/simulator-devices.edn"

def main():
    devs = edn.loads(open(SRC, encoding="utf-8").read())
    out = []
    for d in devs:
        out.append({
            "id": d.get(":id"), "name": d.get(":name"), "num_qubits": d.get(":num-qubits"),
            "type": d.get(":type"), "technology": d.get(":technology"),
            "native_gates": sorted(jsonable(d.get(":native-gates") or [])),
            "noise_model": jsonable(d.get(":noise-model") or {}),
        })
    path = os.path.join(HERE, "device_profiles.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_device_profiles.py",
                   "reference": "resources/simulator-devices.edn", "devices": out}, f, indent=1)
    print(len(out), "devices ->", path)
    for d in out:
        print(" ", d["id"], d["num_qubits"], list((d["noise_model"].get(":gate-noise") or {}).keys()))

if __name__ == "__main__":
    main()
