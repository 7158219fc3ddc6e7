"""python tools/watchdog.py SECONDS script.py [args...]: run a script; after SECONDS dump every thread's Python
traceback to stderr and exit (a hung GPU run then still tells where it hung)."""
import faulthandler, runpy, sys
secs = float(sys.argv[1])
faulthandler.dump_traceback_later(secs, exit=True)
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
