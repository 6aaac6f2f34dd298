"""Operator API kept from the reference: pde_opt/numerics/equations/base_eq.py:11-51."""
from abc import ABC, abstractmethod


class BaseEquation(ABC):
    @abstractmethod
    def rhs(self, state, t):
        raise NotImplementedError


class TimeSplittingEquation(BaseEquation):
    @abstractmethod
    def A_terms(self, state, t):
        raise NotImplementedError

    @abstractmethod
    def B_terms(self, state, t):
        raise NotImplementedError
