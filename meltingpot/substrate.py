"""`meltingpot.substrate`: the reference's public substrate API, served by the B200 engine (see meltingpot_b200.substrate)."""

from meltingpot_b200.substrate import *  # noqa: F401,F403  pylint: disable=wildcard-import
from meltingpot_b200.substrate import (SUBSTRATES, BatchedSubstrate, BatchedTimeStep, Substrate, SubstrateFactory,  # noqa: F401
                                       build, build_batched, build_from_config, get_config, get_factory,
                                       get_factory_from_config)
