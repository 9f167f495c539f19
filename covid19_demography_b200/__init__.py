"""Importable alias of the `covid19-demography_b200/` package directory (a hyphen cannot be imported).

    import covid19_demography_b200 as cb
    cb.seir_individual.run_model(...)        # drop-in for the reference's seir_individual.run_model
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "covid19-demography_b200")]
__all__ = ["engine", "tables", "history", "seir_individual", "population", "build", "ensemble", "driver_tallies", "sharding", "sweep", "sweep_io"]
