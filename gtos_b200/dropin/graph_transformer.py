"""Drop-in module: put this directory in front of the reference's script directory on sys.path and
``generator/generator.py`` / ``translator/generator.py`` import the B200 implementation unchanged."""
from gtos_b200.graph_transformer import *  # noqa: F401,F403
