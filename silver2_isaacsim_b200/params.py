"""Parameter surface of the hydrodynamics engine.

Mirrors the reference's ``HydrodynamicsBehavior.VARIABLES_TO_EXPOSE``
(/root/reference/src/scripts/physics/hydrodynamics_behavior.py:28-46, README.md:128-141):
the same twelve camelCase names and defaults, the same two-level configuration
(``globals`` applied to every prim, then the first ``parts`` key that is a
case-insensitive substring of the prim name; ``hydrodynamics_behavior.py:72-112``),
and the mapping of those names onto the reference wrapper constructor
(``hydrodynamics_behavior.py:155-169``).

The device-side coefficient record (one per body, or one per part type) is
``COEFF_FIELDS`` -- eleven scalars; water density and gravity are engine globals.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, fields
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# name -> default, in the reference's order (hydrodynamics_behavior.py:28-46)
EXPOSED_VARIABLES: Tuple[Tuple[str, float], ...] = (
    ("waterDensity", 1025.0),
    ("gravity", 9.81),
    ("xDimension", 1.0),
    ("yDimension", 1.0),
    ("zDimension", 1.0),
    ("linearDragCoefficient", 1.2),
    ("angularDragCoefficient", 0.8),
    ("linearDamping", 300.0),
    ("angularDamping", 150.0),
    ("linearAddedMassCoefficient", 0.05),
    ("angularAddedMassCoefficient", 0.02),
    ("liftCoefficient", 1.0),
)
EXPOSED_DEFAULTS: Dict[str, float] = dict(EXPOSED_VARIABLES)

# Device coefficient record, per body or per part type (11 scalars).
COEFF_FIELDS: Tuple[str, ...] = (
    "xDimension", "yDimension", "zDimension",
    "linearDragCoefficient", "angularDragCoefficient",
    "linearDamping", "angularDamping",
    "linearAddedMassCoefficient", "angularAddedMassCoefficient",
    "liftCoefficient", "mass",
)
N_COEFF = len(COEFF_FIELDS)

# Reference wrapper ctor order (numba_hydrodynamics_wrapper.py:9-10) expressed in
# exposed-variable names (hydrodynamics_behavior.py:155-169).
CTOR_ORDER: Tuple[str, ...] = (
    "xDimension", "yDimension", "zDimension",
    "linearDragCoefficient", "angularDragCoefficient",
    "linearDamping", "angularDamping",
    "waterDensity", "gravity",
    "linearAddedMassCoefficient", "angularAddedMassCoefficient",
    "liftCoefficient",
)

DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                                   "hydrodynamics_config.json")

# Slot order of the 19 scripted SILVER2 bodies.
HEXAPOD_SLOTS: Tuple[str, ...] = (
    ("Body",) + tuple(f"Coxa_{i}" for i in range(6)) + tuple(f"Femur_{i}" for i in range(6))
    + tuple(f"Tibia_{i}" for i in range(6))
)


@dataclass
class HydroParams:
    """The twelve exposed variables of one prim (reference names, reference defaults)."""

    waterDensity: float = 1025.0
    gravity: float = 9.81
    xDimension: float = 1.0
    yDimension: float = 1.0
    zDimension: float = 1.0
    linearDragCoefficient: float = 1.2
    angularDragCoefficient: float = 0.8
    linearDamping: float = 300.0
    angularDamping: float = 150.0
    linearAddedMassCoefficient: float = 0.05
    angularAddedMassCoefficient: float = 0.02
    liftCoefficient: float = 1.0

    def set(self, name: str, value) -> bool:
        """``_set_attr`` (hydrodynamics_behavior.py:114-121): unknown names are rejected."""
        if name not in EXPOSED_DEFAULTS:
            return False
        setattr(self, name, float(value))
        return True

    def ctor_row(self) -> List[float]:
        """Arguments of the reference wrapper ctor, in its positional order."""
        return [float(getattr(self, k)) for k in CTOR_ORDER]

    def coeff_record(self, mass: float) -> List[float]:
        return [float(getattr(self, k)) for k in COEFF_FIELDS[:-1]] + [float(mass)]

    def as_float32(self) -> "HydroParams":
        """USD stores the exposed variables as ``Sdf.ValueTypeNames.Float`` (fp32)."""
        return HydroParams(**{f.name: float(np.float32(getattr(self, f.name))) for f in fields(self)})


def load_config(path: Optional[str] = None) -> dict:
    with open(path or DEFAULT_CONFIG_PATH, "r") as f:
        return json.load(f)


def match_part(prim_name: str, config: dict) -> Optional[str]:
    """Part-category lookup of ``_apply_json_config`` (hydrodynamics_behavior.py:91-101)."""
    name = prim_name.lower()
    part_type = None
    for category in config.get("parts", {}).keys():
        if category.lower() in name:
            part_type = category
            break
    if part_type is None and "body" in name:
        part_type = "body"
    return part_type


def params_for_prim(prim_name: str, config: Optional[dict] = None,
                    base: Optional[HydroParams] = None) -> Tuple[HydroParams, Optional[str]]:
    """Defaults -> ``globals`` -> matching ``parts`` entry (hydrodynamics_behavior.py:72-112)."""
    config = load_config() if config is None else config
    p = HydroParams(**vars(base)) if base is not None else HydroParams()
    for k, v in config.get("globals", {}).items():
        p.set(k, v)
    part = match_part(prim_name, config)
    if part is not None and part in config.get("parts", {}):
        for k, v in config["parts"][part].items():
            p.set(k, v)
    else:
        part = None
    return p, part


def part_table(prim_names: Sequence[str], masses: Optional[Iterable[float]] = None,
               config: Optional[dict] = None):
    """Deduplicated part-type table for a list of prim names (one articulation).

    Returns ``(table (n_types,11) f64, slot_type (len(prim_names),) int32, rho, g)``.
    ``masses``: per-slot masses; default = ``config['masses'][part]``.
    """
    config = load_config() if config is None else config
    rows: List[List[float]] = []
    slot_type: List[int] = []
    index: Dict[Tuple[float, ...], int] = {}
    rho = g = None
    masses = list(masses) if masses is not None else None
    for i, name in enumerate(prim_names):
        p, part = params_for_prim(name, config)
        if rho is None:
            rho, g = p.waterDensity, p.gravity
        elif (rho, g) != (p.waterDensity, p.gravity):
            raise ValueError("waterDensity/gravity must be uniform across one engine")
        if masses is not None:
            m = float(masses[i])
        else:
            m = float(config.get("masses", {}).get(part or "", 1.0))
        rec = tuple(p.coeff_record(m))
        if rec not in index:
            index[rec] = len(rows)
            rows.append(list(rec))
        slot_type.append(index[rec])
    return (np.asarray(rows, dtype=np.float64), np.asarray(slot_type, dtype=np.int32),
            float(rho), float(g))


def hexapod_table(config: Optional[dict] = None):
    """SILVER2: Body + 6x(Coxa, Femur, Tibia) = 19 slots, 4 part types."""
    return part_table(HEXAPOD_SLOTS, None, config)


def coeff_to_ctor_rows(coeff: np.ndarray, rho: float, g: float) -> np.ndarray:
    """(n,11) coefficient records -> (n,12) reference-ctor rows (for oracles / shims)."""
    coeff = np.asarray(coeff, dtype=np.float64)
    n = coeff.shape[0]
    out = np.empty((n, 12), dtype=np.float64)
    out[:, 0:7] = coeff[:, 0:7]
    out[:, 7] = rho
    out[:, 8] = g
    out[:, 9:12] = coeff[:, 7:10]
    return out
