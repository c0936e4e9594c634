python profiles/debug_a8.py 2>&1 | tail -12
compute-sanitizer --tool memcheck python profiles/debug_a8.py 2>&1 | grep -v "^orig\|^val\|^nonempty\|^K1\|^I4\|^two" | head -30
