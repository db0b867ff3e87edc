"""enflow_b200: B200-native hot path of enflow behind the reference's `enflow.flow` / `enflow.nn` API.

    from enflow_b200.nn.egcl import EGCL
    from enflow_b200.nn.argmax import ArgMax
    from enflow_b200.flow.dynamics import LFIntegrator
    from enflow_b200.flow.loss import Alchemical_NLL
    from enflow_b200.data.base import Data, DataLoader

``enflow_b200.compat.install()`` registers the package under the name ``enflow`` so that code written
against the reference (`from enflow.flow.dynamics import LFIntegrator`) runs unchanged.
"""
__version__ = '0.1.0'
