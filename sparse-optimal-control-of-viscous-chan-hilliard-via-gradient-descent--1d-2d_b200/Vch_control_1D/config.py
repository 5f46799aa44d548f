"""Typed parameters of the 1D vCH control problem — drop-in for the reference's 1D `config.py`
(1D/Vch_control_1D/config.py:91-264): same models, defaults, validators, `last_run_config.json` format and prompts.
Host-side only."""
import json
from typing import Optional, Type

from pydantic import BaseModel, Field, ValidationError

try:
    from pydantic import field_validator as _fv

    def _after(name):
        return _fv(name)

    def _peer(info, key, default=None):
        return (info.data or {}).get(key, default)
except ImportError:                       # pydantic v1
    from pydantic import validator as _v1

    def _after(name):
        return _v1(name)

    def _peer(values, key, default=None):
        return values.get(key, default)

DEFAULT_FILE = "last_run_config.json"


class ForwardSolverConfig(BaseModel):
    """Grid, horizon and physics of the 1D forward solve (reference :93-102)."""
    N: int = Field(128, gt=10, description="Number of spatial intervals")
    Lx: float = Field(1.0, gt=0, description="Domain length")
    T: float = Field(1.0, gt=0, description="Total simulation time")
    dt_initial: float = Field(1e-2, gt=0, description="Initial time step size")
    tau: float = Field(0.05, description="Viscosity  parameter for phi-equation")
    gamma: float = Field(10.0, gt=0, description="Relaxation parameter ")
    c1: float = Field(0.75, description="Flory-Huggins convex coefficient")
    c2: float = Field(1.0, description="Concave (quadratic) coefficient")
    kappa: float = Field(0.03 ** 2, ge=0, description="Gradient energy coefficient")

    @_after("c2")
    def _c2_above_c1(cls, v, info):
        c1 = _peer(info, "c1", 0)
        if v <= c1:
            raise ValueError(f"c2 ({v}) must be greater than c1 ({c1})")
        return v


class OptimizationConfig(BaseModel):
    """Cost weights, step cap, iteration cap, control box (reference :115-123)."""
    b1: float = Field(0.3, ge=0, description="Weight for space-time tracking cost")
    b2: float = Field(13.0, ge=0, description="Weight for terminal cost")
    b3: float = Field(0.0019, ge=0, description="Weight for control energy cost")
    kappa_sparsity: float = Field(0.00009, ge=0, description="Sparsity weight for L1 term")
    alpha_max: float = Field(100.0, gt=0, description="Initial step size for line search")
    max_iter: int = Field(1000, gt=10, description="Max number of gradient descent iterations")
    u_min: float = Field(-1.0, description="Lower bound for the control")
    u_max: float = Field(1.0, description="Upper bound for the control")

    @_after("u_max")
    def _box_not_empty(cls, v, info):
        lo = _peer(info, "u_min")
        if lo is not None and v <= lo:
            raise ValueError("u_max must be strictly greater than u_min.")
        return v


class SimulationParameters(BaseModel):
    forward_solver: ForwardSolverConfig = Field(default_factory=ForwardSolverConfig)
    optimization: OptimizationConfig = Field(default_factory=OptimizationConfig)
    last_run_iterations: int = Field(0, description="Number of iterations from the last run.")


def get_yes_no_input(prompt: str) -> bool:
    while True:
        ans = input(f"{prompt} (y/n): ").lower().strip()
        if ans in ("y", "yes"):
            return True
        if ans in ("n", "no"):
            return False
        print("Invalid input. Please enter 'y' or 'n'.")


def save_params(fwd_config, opt_config, iteration_count, filepath: str = DEFAULT_FILE):
    blob = SimulationParameters(forward_solver=fwd_config, optimization=opt_config, last_run_iterations=iteration_count)
    try:
        with open(filepath, "w") as fh:
            fh.write(blob.model_dump_json(indent=4) if hasattr(blob, "model_dump_json") else json.dumps(blob.dict(), indent=4))
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


def _typed_prompt(name, info, correction=False):
    """Ask for one field; Enter keeps the class default; re-ask until the text parses as the field's type."""
    default, kind = info.default, info.annotation
    tag = "(Correction) " if correction else ""
    while True:
        raw = input(f"-> {tag}Enter '{name}' ({info.description}) [default: {default}]: ").strip()
        if not raw:
            return default
        try:
            return kind(raw)
        except (ValueError, TypeError):
            print(f"  [Error] Invalid format. Please enter a value of type '{getattr(kind, '__name__', kind)}'.")


def get_user_input_for_config(config_model: Type[BaseModel], title: str,
                              previous_instance: Optional[BaseModel] = None) -> BaseModel:
    """Interactive editor (the second, effective definition in the reference, :180-265): shows the last run's values
    for reference, prompts with the class defaults, then re-prompts only fields that fail validation."""
    print("\n" + "=" * 60 + f"\n--- {title} ---")
    fields = config_model.model_fields if hasattr(config_model, "model_fields") else config_model.__fields__
    if previous_instance:
        print("For your reference, here are the parameters from the last run:\n" + "." * 50)
        for name in fields:
            print(f"  {name:<15}: {getattr(previous_instance, name)}")
        print("." * 50)
    print("Please provide new parameters below.\nPress Enter to accept the original default value shown in [brackets].\n" + "=" * 60)
    answers = {name: _typed_prompt(name, info) for name, info in fields.items()}
    while True:
        try:
            cfg = config_model(**answers)
            print("\n✓ Configuration accepted and validated.")
            return cfg
        except ValidationError as exc:
            print("\n" + "!" * 60 + "\n🚨 PARAMETER ERROR: Please correct the following value(s):")
            bad = []
            for err in exc.errors():
                bad.append(err["loc"][0])
                print(f"  - {err['loc'][0]}: {err['msg']}")
            print("!" * 60)
            for name in dict.fromkeys(bad):
                answers[name] = _typed_prompt(name, fields[name], correction=True)
