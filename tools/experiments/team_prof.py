"""One launch of the team kernel on config-4 shapes for ncu:  python tools/team_prof.py [B] [T]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.team_check as tc
tc.timing(b=int(sys.argv[1]) if len(sys.argv) > 1 else 256, t=int(sys.argv[2]) if len(sys.argv) > 2 else 300)
