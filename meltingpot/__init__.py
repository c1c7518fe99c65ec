"""`meltingpot` import path of the B200 engine (alias of `meltingpot_b200`).

Existing code written against the reference package (`from meltingpot import substrate`; `substrate.build(name,
roles=...)`, `substrate.get_config(name)`, `SUBSTRATES`, ...; `/root/reference/meltingpot/substrate.py:38-113`) resolves
here to `meltingpot_b200.substrate`. Only the hot path's API surface is aliased (substrates, and the scenario wrapper
class); bots, scenario configs and evaluation utilities of the reference are out of scope.
"""

from meltingpot import substrate  # noqa: F401

__all__ = ['substrate']
