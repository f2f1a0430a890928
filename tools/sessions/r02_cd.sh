#!/bin/bash
cd "$(dirname "$0")/../.."
bash tools/sessions/r02_c.sh
bash tools/sessions/r02_d.sh
