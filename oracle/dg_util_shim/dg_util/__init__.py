"""Minimal stand-in for the un-vendored, un-pinned `dg_util` package.

TEST INFRASTRUCTURE ONLY.  The reference (danielgordon10/vince) imports
`dg_util.python_utils.*` everywhere (requirements.txt:21) but the package is
not installable here (no network).  This shim provides just the symbols the
hot path touches so the *unmodified* reference modules can be imported on CPU
to (a) validate oracle/vince_oracle.py and (b) generate tests/golden vectors.
Semantics are inferred from the reference's own call-site comments
(SURVEY.md Appendix A).  Nothing in vince_b200/ imports this.
"""
