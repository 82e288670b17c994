"""Freeze the per-type parameter tables of the reference's shipped URDFs.

Run in a container that has the reference checkout:
    python tools/freeze_vehicle_tables.py [/root/reference/dronesim/assets]
Writes dronesim_b200/assets/vehicle_tables.json (data extracted from the URDFs by
dronesim_b200.vehicles.parse_urdf; no reference source is copied).
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dronesim_b200.vehicles import parse_urdf  # noqa: E402

SHIPPED = ["tello", "robobee", "hexa_6DOF", "hexa_6DOF_simple"]

if __name__ == "__main__":
    assets = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dronesim/assets"
    out = {n: parse_urdf(os.path.join(assets, n + ".urdf")).to_json() for n in SHIPPED}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dronesim_b200", "assets",
                       "vehicle_tables.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(dst))
    # the one propeller the "advanced" quad model uses (BaseAviary.py:1617: propeller = "mamr-8x4.5", lower-order
    # oblique-flow fit): its 14 coefficients from the reference's database module, as data
    ref_root = os.path.normpath(os.path.join(assets, "..", ".."))
    sys.path.insert(0, ref_root)
    from dronesim.database.propeller_database import Data_section5_ObliqueFlow  # noqa: E402

    props = {"mamr-8x4.5": {"section5_oblique_flow": [float(x) for x in Data_section5_ObliqueFlow["mamr-8x4.5"]],
                            "diameter_in": 8.0}}
    dst2 = os.path.join(os.path.dirname(dst), "propeller_tables.json")
    with open(dst2, "w") as f:
        json.dump(props, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(dst2))
