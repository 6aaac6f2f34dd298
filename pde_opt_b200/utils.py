"""Plugin wiring kept verbatim in behaviour from pde_opt/utils.py:6-53."""


def check_equation_solver_compatibility(solver_type, equation_type):
    """Class-level hasattr check (utils.py:6-31); raises ValueError listing missing attrs."""
    if not hasattr(solver_type, "required_equation_attrs"):
        return
    missing = [a for a in solver_type.required_equation_attrs if not hasattr(equation_type, a)]
    if missing:
        raise ValueError(
            f"Equation type {equation_type.__name__} is missing required "
            f"attributes for solver {solver_type.__name__}: {missing}"
        )


def prepare_solver_params(solver_type, solver_parameters, equation):
    """Copy the solver's required attributes from the equation by name (utils.py:34-53)."""
    full = dict(solver_parameters)
    if hasattr(solver_type, "required_equation_attrs"):
        for name in solver_type.required_equation_attrs:
            full[name] = getattr(equation, name)
    return full
