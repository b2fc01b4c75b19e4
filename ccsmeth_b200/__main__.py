"""``python -m ccsmeth_b200 <sub-command> ...`` -- the two sub-commands of the reference's ``ccsmeth`` entry point
(ccsmeth/ccsmeth.py:68-110) that lie on the hot path: ``call_mods`` and ``call_freqb``, with the reference's flags
(tests/test_cli_cpu.py).  Under torchrun every rank runs the same command line on its own GPU."""
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    subs = ("call_mods", "call_freqb")
    if not argv or argv[0] in ("-h", "--help") or argv[0] not in subs:
        sys.stderr.write("usage: python -m ccsmeth_b200 {%s} [flags]\n"
                         "The other ccsmeth sub-commands (call_hifi, align_hifi, call_freqt, extract, train, trainm) are "
                         "outside the B200 hot path; use the reference for them.\n" % ",".join(subs))
        return 0 if argv and argv[0] in ("-h", "--help") else 2
    if argv[0] == "call_mods":
        from .call_mods import main as run
    else:
        from .call_freqb import main as run
    return run(argv[1:])


if __name__ == "__main__":
    sys.exit(main())
