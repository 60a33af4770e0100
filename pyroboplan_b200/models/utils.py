"""Asset lookup for the bundled models (cf. /root/reference/src/pyroboplan/models/utils.py:7-17)."""

import json
import os


def get_example_models_folder():
    """Folder holding the bundled model descriptions (``*.json``)."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def load_description(name):
    with open(os.path.join(get_example_models_folder(), f"{name}.json")) as f:
        return json.load(f)
