"""Drop-in check of the two command lines on the path: every flag of the reference's `call_mods` and `call_freqb`
sub-commands (fixture cli_flags.json, captured from the parser object the reference's main() builds -- scripts/
gen_golden.py gen_cli) exists in the ccsmeth_b200 parsers with the same aliases, default, type, choices and arity.
Flags the B200 CLIs add on top are listed explicitly."""
import argparse
import json
import os

import pytest

from ccsmeth_b200 import call_freqb, call_mods
from tests.conftest import GOLDEN

REF = json.load(open(os.path.join(GOLDEN, "cli_flags.json")))

PARSERS = {"call_mods": call_mods.build_parser, "call_freqb": call_freqb.build_parser}
# additions of this implementation (documented in INTEGRATION.md); everything else must come from the reference
EXTRA = {"call_mods": {"h0", "device_batch", "precision", "bam_compress"},
         "call_freqb": {"h0"}}


def _actions(parser):
    return {a.dest: a for a in parser._actions if not isinstance(a, argparse._HelpAction)}


@pytest.mark.parametrize("cmd", sorted(PARSERS))
def test_every_reference_flag_is_accepted_with_the_same_meaning(cmd):
    mine = _actions(PARSERS[cmd]())
    for ref in REF[cmd]:
        assert ref["dest"] in mine, "%s: missing %s" % (cmd, ref["flags"])
        a = mine[ref["dest"]]
        assert set(ref["flags"]) <= set(a.option_strings), (cmd, ref["flags"], a.option_strings)
        assert a.default == ref["default"], (cmd, ref["dest"], a.default, ref["default"])
        assert getattr(a.type, "__name__", None) == ref["type"], (cmd, ref["dest"])
        assert type(a).__name__ == ref["action"], (cmd, ref["dest"])
        assert bool(a.required) == ref["required"], (cmd, ref["dest"])
        if ref["choices"] is not None:
            assert a.choices is not None and set(ref["choices"]) == set(a.choices), (cmd, ref["dest"])


@pytest.mark.parametrize("cmd", sorted(PARSERS))
def test_additions_are_the_documented_ones(cmd):
    mine = set(_actions(PARSERS[cmd]()))
    assert mine - {r["dest"] for r in REF[cmd]} == EXTRA[cmd]


def test_module_entry_point_dispatches_the_two_sub_commands(capsys):
    from ccsmeth_b200.__main__ import main
    assert main([]) == 2 and main(["train"]) == 2            # off-path sub-commands are refused, not ignored
    with pytest.raises(SystemExit) as e:
        main(["call_freqb", "--help"])
    assert e.value.code == 0 and "--aggre_model" in capsys.readouterr().out
    with pytest.raises(SystemExit):                           # the reference's required flags stay required
        main(["call_mods", "-i", "x.bam"])
