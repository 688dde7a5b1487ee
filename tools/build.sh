#!/bin/sh
cd /root/repo && python __graft_entry__.py | tail -1
