# Per-task run script TEMPLATE of the artefact-compatibility mode and of the host pool's
# python_path / conda_env workers.  Same role, same three substitution fields and same
# observable behaviour as the reference's template
# (/root/reference/covalent_ssh_plugin/exec.py:12-46, formatted at ssh.py:164-169):
#   - reads a cloudpickled (fn, args, kwargs) tuple from remote_function_file;
#   - cloudpickle missing -> ships (None, ImportError) with the stdlib pickle and exits 1;
#   - runs the electron inside current_remote_workdir (created on demand) and always
#     returns to the directory it started in;
#   - an electron that raises still exits 0: the exception travels in-band as the second
#     element of the (result, exception) tuple written to remote_result_file.
# The file must stay str.format-safe: the three fields below are its only braces.
import os
import sys

RESULT_FILE = os.path.realpath("{remote_result_file}")
FUNCTION_FILE = os.path.realpath("{remote_function_file}")
WORKDIR = os.path.realpath("{current_remote_workdir}")

try:
    import cloudpickle as codec
except Exception as import_error:
    import pickle as codec

    with open(RESULT_FILE, "wb") as sink:
        codec.dump((None, import_error), sink)
    sys.exit(1)

with open(FUNCTION_FILE, "rb") as source:
    electron, positional, named = codec.load(source)

outcome = None
failure = None
started_in = os.getcwd()
try:
    os.makedirs(WORKDIR, exist_ok=True)
    os.chdir(WORKDIR)
    outcome = electron(*positional, **named)
except Exception as raised:
    failure = raised
finally:
    os.chdir(started_in)

with open(RESULT_FILE, "wb") as sink:
    codec.dump((outcome, failure), sink)
