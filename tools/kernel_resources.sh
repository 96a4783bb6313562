#!/bin/bash
# registers / shared memory / spills of every kernel of the built library (cuobjdump -res-usage), one line each
cuobjdump -res-usage "${1:-web-audio-api-rs_b200/libwae_b200.so}" 2>/dev/null | awk '/Function/ {name=$2} /REG:/ {print name, $0}' | sed 's/:$//' | while read -r name rest; do
  echo "$(echo "${name%:}" | c++filt | sed 's/(.*//') $rest"
done
