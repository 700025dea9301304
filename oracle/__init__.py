"""TEST INFRASTRUCTURE: CPU restatement of farr/mcmc-ocaml's sampling-and-
evidence path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package."""
