"""B200-native daily agent-stepping path of the COVID19-Demography SEIR simulator.

This directory is both
  * a sys.path entry holding the drop-in module `seir_individual` (same name, same two entry points
    as /root/reference/seir_individual.py: `run_complete_simulation`, `run_model`), and
  * the package imported as `covid19_demography_b200` (see the alias package next to it).
"""
