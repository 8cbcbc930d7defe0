import faulthandler, sys, os
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
g.smoke()
