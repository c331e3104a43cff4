"""Names-only stand-in for `jsonpickle` (a dependency of the reference's data/huggingface_utils.py that this image does not
ship), so that the reference's own serialize / huggingface_utils modules can be imported by the build-container tests.
TEST INFRASTRUCTURE.  `encode` / `decode` are plain JSON: identical strings to jsonpickle's for the JSON-representable info
dicts the tests use (jsonpickle only differs for objects that need type tags)."""
import json


def encode(value, *args, **kwargs) -> str:
    return json.dumps(value)


def decode(string, *args, **kwargs):
    return json.loads(string)
