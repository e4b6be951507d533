"""The from-YAML front end of the mechanism pipeline (tools/yaml_to_mechanism.py).

Known answers: the reference ships every mechanism both as Cantera YAML and as the CEPTR-generated mechanism.H
(PP/Mechanisms/<mech>/).  The table made from the YAML must equal the table extracted from the header (the committed
pelec_b200/mechanisms/<mech>.json the parity tests pin) string for string - species, molecular weights, NASA-7 forms,
Arrhenius / Troe literals, stoichiometry and the reference's reaction order - so the CUDA / C code generated from either
is the same code.  Needs /root/reference (build container); the unit-handling and rejection tests do not."""
import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import yaml_to_mechanism as Y  # noqa: E402
from pelec_b200.mechgen import gen  # noqa: E402

MECH_DIR = os.path.join(os.environ.get("PELE_REFERENCE", "/root/reference"), "Submodules/PelePhysics/Mechanisms")
SHIPPED = ["LiDryer", "drm19", "dodecane_lu", "BurkeDryer", "FFCM1_Red"]


def strip_comments(text):
    return re.sub(r"//[^\n]*", "", text)


@pytest.mark.parametrize("mech", SHIPPED)
def test_yaml_table_equals_header_table(mech):
    path = os.path.join(MECH_DIR, mech, "mechanism.yaml")
    if not os.path.exists(path):
        pytest.skip("reference mechanisms not present")
    committed = os.path.join(gen.MECH_JSON_DIR, mech + ".json")
    a = Y.convert(path, mech, transport_from=committed)
    b = json.load(open(committed))
    Y.compare_tables(a, b, rtol=0.0)  # structure, order and every number
    for x, y in zip(a["reactions"], b["reactions"]):
        assert {k: v for k, v in x.items() if k != "eq"} == {k: v for k, v in y.items() if k != "eq"}, y["eq"]
    for key in ("species", "mw", "imw", "thermo", "Ru", "refC_num", "refC_den", "T_newton", "transport"):
        assert a[key] == b[key], key
    # the same table gives the same code (comments carry the equation as each route spells it)
    for flavour in ("cuda", "c"):
        assert strip_comments(gen.Emitter(a, flavour).emit()) == strip_comments(gen.Emitter(b, flavour).emit())


TOY = """
units: {length: %(length)s, quantity: %(quantity)s, activation-energy: %(eunit)s}
phases:
- name: gas
  species: [H2, H, NO]
species:
- name: H2
  composition: {H: 2}
  thermo: {model: NASA7, temperature-ranges: [200.0, 1000.0, 3500.0], data: [[2.34, 7.9e-03, -1.9e-05, 2.0e-08, -7.3e-12, -917.9, 0.683], [3.33, -4.9e-05, 4.9e-07, -1.7e-10, 2.0e-14, -950.1, -3.20]]}
- name: H
  composition: {H: 1}
  thermo: {model: NASA7, temperature-ranges: [200.0, 1000.0, 3500.0], data: [[2.5, 0.0, 0.0, 0.0, 0.0, 25473.6, -0.4466], [2.5, 0.0, 0.0, 0.0, 0.0, 25473.6, -0.4466]]}
- name: NO
  composition: {N: 1, O: 1}
  thermo: {model: NASA7, temperature-ranges: [200.0, 1000.0, 6000.0], data: [[4.2, -4.6e-03, 1.1e-05, -9.3e-09, 2.8e-12, 9844.6, 2.28], [3.26, 1.19e-03, -4.3e-07, 6.9e-11, -4.0e-15, 9920.9, 6.37]]}
reactions:
- equation: 2 H + M <=> H2 + M
  type: three-body
  rate-constant: {A: %(A3)s, b: -1.0, Ea: 0.0}
  efficiencies: {H2: 2.5}
- equation: H2 + NO <=> 2 H + NO
  rate-constant: {A: %(A2)s, b: 0.5, Ea: %(Ea)s}
- equation: 2 H (+M) <=> H2 (+M)
  type: falloff
  low-P-rate-constant: {A: %(A3)s, b: -1.0, Ea: 0.0}
  high-P-rate-constant: {A: %(A2)s, b: 0.0, Ea: %(Ea)s}
  Troe: {A: 0.5, T3: 1.0e-30, T1: 1.0e+30}
%(extra)s
"""


def toy(tmp_path, **kw):
    p = tmp_path / "toy.yaml"
    args = dict(length="cm", quantity="mol", eunit="cal/mol", A3="1.0e+18", A2="1.0e+13", Ea="1000.0", extra="")
    args.update(kw)
    p.write_text(TOY % args)
    return str(p)


def test_units_and_ordering(tmp_path):
    a = Y.convert(toy(tmp_path), "toy")
    # the same mechanism in SI-ish units: m, kmol, J/mol
    b = Y.convert(toy(tmp_path, length="m", quantity="kmol", eunit="J/mol", A3="1.0e+12", A2="1.0e+10", Ea="4184.0"), "toy")
    Y.compare_tables(a, b, rtol=1e-14)
    assert a["species"] == ["H2", "H", "NO"]  # NO stays a name (YAML 1.1 would read a boolean)
    assert [r["type"] for r in a["reactions"]] == ["falloff", "three_body", "elementary"]  # the reference's order
    tb = a["reactions"][1]
    assert float(tb["A"]) == 1.0e18 * 1e-12 and tb["eff"] == [[0, "1.5"]] and tb["refC"] == 1
    # 'H2 + NO <=> 2 H + NO': 2 -> 3 molecules with one species on both sides = explicit collider (Cantera 3): NO leaves the
    # stoichiometry but stays in the concentration products, and the reaction sorts with the third-body group
    el = a["reactions"][2]
    assert el["reactants"] == [0, 2] and el["products"] == [1, 1, 2]
    assert el["wdot"] == [[0, "-1"], [1, "2.000000"]] and el["dG"] == [["1", 0], ["-2.000000", 1]] and el["refC"] == -1
    assert abs(float(el["Ea"]) - 1000.0 / 1.98721558317399615845) < 1e-9
    # H: the two NASA rows coincide -> one range
    assert a["thermo"]["gibbs"][1]["tmid"] is None and a["thermo"]["gibbs"][0]["tmid"] == 1000.0
    # the table feeds the generator
    assert "production_rate" in gen.Emitter(a, "cuda").emit()


@pytest.mark.parametrize("extra", [
    "- equation: H2 <=> 2 H\n  type: pressure-dependent-Arrhenius\n  rate-constants: [{P: 1 atm, A: 1.0, b: 0.0, Ea: 0.0}]",
    "- equation: H2 + M <=> 2 H + M\n  type: three-body\n  rate-constant: {A: 1.0, b: 0.0, Ea: 0.0}\n  default-efficiency: 0.0",
    "- equation: 0.5 H2 <=> H\n  rate-constant: {A: 1.0, b: 0.0, Ea: 0.0}",
    "- equation: H2 <=> 2 H\n  rate-constant: {A: 1.0, b: 0.0, Ea: 0.0}\n  orders: {H2: 0.5}",
])
def test_unsupported_input_is_rejected(tmp_path, extra):
    with pytest.raises((NotImplementedError, ValueError)):
        Y.convert(toy(tmp_path, extra=extra), "toy")


@pytest.mark.parametrize("mech", ["propane_fc", "LuDME", "grimech30", "C1-C2-NO"])
def test_two_front_ends_agree_on_other_reference_mechanisms(mech):
    """Mechanisms that are not in the default library: the two independent routes (header extractor, YAML converter) must
    give the same table.  propane_fc has irreversible third-body / falloff reactions (printed as '... * (0.0)' in the header),
    LuDME a reaction with delta n = 2 (doubled parentheses) - the two print forms the cross-check found the extractor had
    misread; grimech30 / C1-C2-NO have the species NO (a YAML 1.1 boolean)."""
    import extract_mechanism as X
    path = os.path.join(MECH_DIR, mech, "mechanism.yaml")
    if not os.path.exists(path):
        pytest.skip("reference mechanisms not present")
    a = Y.convert(path, mech)
    b = X.extract(mech)
    Y.compare_tables(a, b, rtol=1e-12)
    assert any(not r["reversible"] for r in b["reactions"]) or mech != "propane_fc"
    assert any(abs(r["refC"]) == 2 for r in b["reactions"]) or mech != "LuDME"
