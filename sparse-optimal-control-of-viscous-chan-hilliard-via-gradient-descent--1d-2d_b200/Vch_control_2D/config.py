"""Typed parameters of the 2D vCH control problem — drop-in for the reference's 2D `config.py`
(2D/Vch_control_2D/config.py:83-190): same model names, fields, defaults, validators, JSON file and prompt helpers,
so configs saved by either code load in the other.  Host-side only; nothing here touches the GPU.
"""
import json
from typing import Any, Dict, Optional, Type

from pydantic import BaseModel, Field, ValidationError

try:                                    # pydantic v2
    from pydantic import field_validator as _fv

    def _after(field):
        return _fv(field)

    def _other(info, name, default=None):
        return (info.data or {}).get(name, default)
except ImportError:                     # pydantic v1
    from pydantic import validator as _v1

    def _after(field):
        return _v1(field)

    def _other(values, name, default=None):
        return values.get(name, default)

DEFAULT_FILE = "last_run_config_2d.json"


class ForwardSolverConfig(BaseModel):
    """Grid, horizon and physics of the forward solve (reference defaults, config.py:103-113)."""
    Nx: int = Field(128, gt=10, description="Number of spatial intervals in x")
    Ny: int = Field(128, gt=10, description="Number of spatial intervals in y")
    Lx: float = Field(1.0, gt=0, description="Domain length in x")
    Ly: float = Field(1.0, gt=0, description="Domain length in y")
    T: float = Field(1.0, gt=0, description="Total simulation time")
    dt_initial: float = Field(1e-2, gt=0, description="Initial time step size")
    tau: float = Field(0.05, description="viscosity parameter for phi-equation")
    gamma: float = Field(10.0, gt=0, description="Relaxation parameter ")
    c1: float = Field(0.75, description="Flory–Huggins convex coefficient")
    c2: float = Field(1.0, description="Concave (quadratic) coefficient")
    kappa: float = Field(0.01 ** 2, ge=0, description="Gradient energy coefficient")

    @_after("c2")
    def _c2_above_c1(cls, v, info):
        c1 = _other(info, "c1", 0.0)
        if v <= c1:
            raise ValueError(f"c2 ({v}) must be greater than c1 ({c1})")
        return v


class OptimizationConfig(BaseModel):
    """Cost weights, step-size cap, iteration cap and control box (reference defaults, config.py:137-144)."""
    b1: float = Field(5.0, ge=0, description="Weight for space-time tracking cost")
    b2: float = Field(10.0, ge=0, description="Weight for terminal cost")
    b3: float = Field(0.0001, ge=0, description="Weight for control energy cost")
    kappa_sparsity: float = Field(1e-4, ge=0, description="Sparsity weight for L1 term")
    alpha_max: float = Field(50.0, gt=0, description="Initial step size for line search")
    max_iter: int = Field(500, gt=10, description="Max number of gradient descent iterations")
    u_min: float = Field(-1.0, description="Lower bound for the control")
    u_max: float = Field(1.0, description="Upper bound for the control")

    @_after("u_max")
    def _box_not_empty(cls, v, info):
        lo = _other(info, "u_min")
        if lo is not None and v <= lo:
            raise ValueError("u_max must be strictly greater than u_min.")
        return v


class SimulationParameters(BaseModel):
    forward_solver: ForwardSolverConfig = Field(default_factory=ForwardSolverConfig)
    optimization: OptimizationConfig = Field(default_factory=OptimizationConfig)
    last_run_iterations: int = Field(0, description="Number of iterations from the last run.")


def _as_dict(m: BaseModel) -> Dict[str, Any]:
    return m.model_dump() if hasattr(m, "model_dump") else m.dict()


def _as_json(m: BaseModel, indent: int = 4) -> str:
    return m.model_dump_json(indent=indent) if hasattr(m, "model_dump_json") else json.dumps(m.dict(), indent=indent)


def _fields(cls: Type[BaseModel]) -> Dict[str, Any]:
    return cls.model_fields if hasattr(cls, "model_fields") else cls.__fields__


def save_params(fwd_config, opt_config, iteration_count, filepath: str = DEFAULT_FILE) -> None:
    blob = SimulationParameters(forward_solver=fwd_config, optimization=opt_config, last_run_iterations=iteration_count)
    try:
        with open(filepath, "w") as fh:
            fh.write(_as_json(blob))
        print(f"\n✅ Configuration saved to '{filepath}' for your next session.")
    except IOError as exc:
        print(f"\n[Warning] Could not save configuration file: {exc}")


def load_params(filepath: str = DEFAULT_FILE) -> SimulationParameters:
    try:
        with open(filepath) as fh:
            blob = SimulationParameters(**json.load(fh))
        print(f"✅ Loaded previous configuration from '{filepath}'.")
        return blob
    except (FileNotFoundError, ValidationError, json.JSONDecodeError):
        print("No valid previous configuration found. Using default parameters.")
        return SimulationParameters()


def get_yes_no_input(prompt: str) -> bool:
    while True:
        ans = input(f"{prompt} (y/n): ").strip().lower()
        if ans in ("y", "yes"):
            return True
        if ans in ("n", "no"):
            return False
        print("Invalid input. Please enter 'y' or 'n'.")


def _ask(name: str, info: Any, correction: bool = False):
    default = getattr(info, "default", None)
    desc = getattr(info, "description", None) or getattr(getattr(info, "field_info", None), "description", "") or ""
    tag = "(Correction) " if correction else ""
    raw = input(f"-> {tag}Enter '{name}' ({desc}) [default: {default}]: ").strip()
    return default if raw == "" else raw


def get_user_input_for_config(config_model: Type[BaseModel], title: str,
                              previous_instance: Optional[BaseModel] = None) -> BaseModel:
    """Prompt for every field (Enter keeps the default), then re-prompt only the fields Pydantic rejects."""
    print("\n" + "=" * 60 + f"\n--- {title} ---")
    if previous_instance is not None:
        print("For your reference, here are the parameters from the last run:\n" + "." * 50)
        for k, v in _as_dict(previous_instance).items():
            print(f"  {k:<15}: {v}")
        print("." * 50)
    print("Please provide new parameters below.\nPress Enter to accept the original default value shown in [brackets].\n" + "=" * 60)
    fields = _fields(config_model)
    answers = {name: _ask(name, info) for name, info in fields.items()}
    while True:
        try:
            cfg = config_model(**answers)
            print("\n✓ Configuration accepted and validated.")
            return cfg
        except ValidationError as exc:
            print("\n" + "!" * 60 + "\n🚨 PARAMETER ERROR: Please correct the following value(s):")
            bad = []
            for err in exc.errors():
                if err.get("loc"):
                    bad.append(err["loc"][0])
                    print(f"  - {err['loc'][0]}: {err['msg']}")
            print("!" * 60)
            for name in dict.fromkeys(bad):
                answers[name] = _ask(name, fields[name], correction=True)
