// ORACLE shim: upstream splits the API over several headers; the restatement keeps one.
#pragma once
#include <simplicial_arrangement/simplicial_arrangement.h>
